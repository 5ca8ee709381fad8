// standalone probe of the MN-major UMMA path (debug aid; not part of the library)
#include <cstdio>
#include <vector>
#include <cstdlib>
#include "../gigl_b200/csrc/gemm_tn_tcgen05.cu"
int gigl_fail(gigl_ctx* ctx, int code, const std::string& msg) { printf("FAIL %d %s\n", code, msg.c_str()); return code; }
int gigl_cuda_fail(gigl_ctx* ctx, cudaError_t e, const char* what) { printf("CUDA FAIL %s %s\n", what, cudaGetErrorString(e)); return -2; }
int gigl_scratch(gigl_ctx* ctx, int slot, size_t bytes, void** out) {
    if (ctx->scratch_bytes[slot] < bytes) { cudaMalloc(&ctx->scratch[slot], bytes); cudaMemset(ctx->scratch[slot], 0x7f, bytes); ctx->scratch_bytes[slot] = bytes; }
    *out = ctx->scratch[slot]; return 0; }
int main(int argc, char** argv) {
    gigl_ctx ctx; cudaStreamCreate(&ctx.stream);
    int R = 16, M = 128, N = 64;
    std::vector<float> G(R * M, 0.f), A(R * N, 1.f), Z(R * (M > N ? M : N), 0.f), C(M * N, -5.f);
    int gr = argc > 1 ? atoi(argv[1]) : 0, gm = argc > 2 ? atoi(argv[2]) : 0;
    G[gr * M + gm] = 1.f;
    float *dG, *dA, *dZ, *dC;
    cudaMalloc(&dG, G.size() * 4); cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dZ, Z.size() * 4); cudaMalloc(&dC, C.size() * 4);
    cudaMemcpy(dG, G.data(), G.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dZ, Z.data(), Z.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dC, C.data(), C.size() * 4, cudaMemcpyHostToDevice);
    int rc = linear_tn_tc_launch(&ctx, R, M, N, dG, dZ, M, dA, dZ, N, dC, N, N, N, nullptr, 0, 0, 0);
    cudaError_t e = cudaDeviceSynchronize();
    printf("rc %d sync %s\n", rc, cudaGetErrorString(e));
    cudaMemcpy(C.data(), dC, C.size() * 4, cudaMemcpyDeviceToHost);
    std::vector<float> P(128 * 64);
    cudaMemcpy(P.data(), ctx.scratch[GIGL_SLOT_WORK], P.size() * 4, cudaMemcpyDeviceToHost);
    int nzc = 0, nzp = 0;
    for (int i = 0; i < M * N; ++i) { if (C[i] != 0.f) { if (nzc < 8) printf("C[%d,%d]=%g\n", i / N, i % N, C[i]); nzc++; } }
    for (int i = 0; i < 128 * 64; ++i) { if (P[i] != 0.f) { if (nzp < 8) printf("P[%d,%d]=%g\n", i / 64, i % 64, P[i]); nzp++; } }
    printf("nonzero C %d, nonzero partial %d\n", nzc, nzp);
    return 0;
}
