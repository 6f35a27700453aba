// Timeline probe for the attention backward kernel: stamps clock64() at the pipeline hand-offs of ONE CTA and prints
// the per-iteration deltas.  Build + run on the box:
//   nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 --use_fast_math -DB200_ATTN_TRACE -Ifms_fsdp_b200/csrc \
//        scripts/attn_trace.cu -o /tmp/attn_trace -lcuda && /tmp/attn_trace
#include "../fms_fsdp_b200/csrc/attention_sm100.cu"
#include <cstdio>
#include <cstdlib>
#include <vector>

int main(int argc, char** argv) {
  const int B = 2, S = 4096, H = 32, KVH = 32, HD = 128;
  const int cta = argc > 1 ? atoi(argv[1]) : 0;
  const size_t W = (size_t)(H + 2 * KVH) * HD, M = (size_t)B * S;
  std::vector<__nv_bfloat16> h(M * W);
  srand(1);
  for (auto& v : h) v = __float2bfloat16((rand() % 2001 - 1000) / 1000.f);
  __nv_bfloat16 *qkv, *o, *dout, *dqkv;
  float *lse, *delta;
  cudaMalloc(&qkv, M * W * 2); cudaMalloc(&dqkv, M * W * 2);
  cudaMalloc(&o, M * H * HD * 2); cudaMalloc(&dout, M * H * HD * 2);
  cudaMalloc(&lse, (size_t)B * H * S * 4); cudaMalloc(&delta, (size_t)2 * B * H * (S + 128) * 4);
  cudaMemcpy(qkv, h.data(), M * W * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dout, h.data(), M * H * HD * 2, cudaMemcpyHostToDevice);
  b200_attn_set_poly(argc > 2 ? atoi(argv[2]) : 2, argc > 2 ? atoi(argv[2]) : 2);
  int rc = b200_attn_fwd(qkv, o, lse, B, S, H, KVH, HD, 0.0884f, 0);
  printf("fwd rc %d\n", rc);
  long long* tr;
  cudaMalloc(&tr, 64 * 16 * 8);
  for (int mode = 0; mode < 2; ++mode) {   // mode 0 = dK/dV kernel, 1 = dQ kernel
    cudaMemset(tr, 0, 64 * 16 * 8);
    cudaMemcpyToSymbol(b200::g_attn_trace, &tr, sizeof(tr));
    cudaMemcpyToSymbol(b200::g_attn_trace_cta, &cta, sizeof(int));
    cudaMemcpyToSymbol(b200::g_attn_trace_mode, &mode, sizeof(int));
    rc = b200_attn_bwd(dout, qkv, o, lse, dqkv, delta, B, S, H, KVH, HD, 0.0884f, nullptr, 0);
    cudaError_t e = cudaDeviceSynchronize();
    printf("bwd rc %d sync %s\n", rc, cudaGetErrorString(e));
  std::vector<long long> t(64 * 16);
  cudaMemcpy(t.data(), tr, 64 * 16 * 8, cudaMemcpyDeviceToHost);
  long long t0 = t[0];
  printf("v3 kernel mode %d.  slots 0-3 mma: p_full seen | acc1+S(k+1) issued | ds_full seen | acc2+dP(k+1) issued || wg0: 4 E wait, 5 s_full, 6 E done(p arrive), 7 dp_full, 8 D done, 9 S loaded, 10 exp done, 11 dP loaded\n", mode);
  for (int it = 0; it < 24; ++it) {
    printf("%2d |", it);
    for (int s = 0; s < 12; ++s) printf(" %7lld", t[it * 16 + s] ? t[it * 16 + s] - t0 : -1);
    printf("\n");
  }
  }
  return 0;
}
