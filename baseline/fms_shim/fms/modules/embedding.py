import torch
import torch.nn as nn


class WordEmbedding(nn.Module):
    def __init__(self, vocab_size, emb_dim, padding_idx=None, reversible=True, tie_weights=False, bias=False):
        super().__init__()
        self.vocab_size, self.emb_dim = vocab_size, emb_dim
        self.emb = nn.Embedding(vocab_size, emb_dim)
        self.head = nn.Linear(emb_dim, vocab_size, bias=False)
        if tie_weights:
            self.head.weight = self.emb.weight

    def reset_parameters(self):
        nn.init.trunc_normal_(self.emb.weight, mean=0.0, std=self.emb_dim ** -0.5)
        nn.init.trunc_normal_(self.head.weight, mean=0.0, std=self.emb_dim ** -0.5)

    def forward(self, inp, reverse=False):
        return self.head(inp) if reverse else self.emb(inp)
