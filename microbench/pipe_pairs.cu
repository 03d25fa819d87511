// pipe_pairs.cu -- which of the packed-16-bit / half2 / integer instructions share an issue pipe
// on sm_100a?  For each (A, B) pair runs NA ops of A and NB ops of B per step on independent
// register chains, one 1024-thread block per SM (exact per-SM accounting with clock64), and
// prints thread-instructions / clock / SM.  64 = one 16-lane pipe per SM sub-partition saturated;
// 128 = two pipes running concurrently (issue limit).  Standalone; not part of the product.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
    fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)

enum { VIADDMNMX, VIMNMX, VIMNMX3, VIADD16, IMAD, HFMA2R, HADD2, VHMNMX, LOP3, IADD3, SHF, PRMT, FFMA, FMNMX, VIADDMNMX_IMM, NOPS };
static const char *names[] = {"VIADDMNMX.S16x2", "VIMNMX.S16x2", "VIMNMX3.S16x2", "VIADD.16x2", "IMAD", "HFMA2.RELU",
                              "HADD2", "VHMNMX(f16x2 max)", "LOP3", "IADD3", "SHF", "PRMT", "FFMA", "FMNMX",
                              "VIADDMNMX.S16x2 (immediate addend)"};

template <int OP>
__device__ __forceinline__ void step(unsigned &v, unsigned &w, unsigned k1, unsigned k2, unsigned one)
{
    if (OP == VIADDMNMX) v = __viaddmax_s16x2(v, k1, w);
    if (OP == VIADDMNMX_IMM) v = __viaddmax_s16x2(v, 0xfffcfffcu, w);
    if (OP == VIMNMX)  { unsigned t = __vmaxs2(v, w); w = v; v = t; }             // 1 VIMNMX (+ renaming)
    if (OP == VIMNMX3)   v = __vimax3_s16x2(v, k1, w);
    if (OP == VIADD16)   v = __vadd2(v, k1);
    if (OP == IMAD)      asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(v) : "r"(v), "r"(one), "r"(k1));
    if (OP == HFMA2R)    asm volatile("fma.rn.relu.f16x2 %0, %1, %2, %3;" : "=r"(v) : "r"(v), "r"(k2), "r"(k1));
    if (OP == HADD2)     asm volatile("add.rn.f16x2 %0, %1, %2;" : "=r"(v) : "r"(v), "r"(k1));
    if (OP == VHMNMX)  { unsigned t; asm volatile("max.f16x2 %0, %1, %2;" : "=r"(t) : "r"(v), "r"(w)); w = v; v = t; }
    if (OP == LOP3)      asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(v) : "r"(v), "r"(k1), "r"(w));
    if (OP == IADD3)     asm volatile("add.u32 %0, %1, %2;" : "=r"(v) : "r"(v), "r"(w));
    if (OP == SHF)       asm volatile("shf.l.wrap.b32 %0, %1, %2, %3;" : "=r"(v) : "r"(v), "r"(w), "r"(k1));
    if (OP == PRMT)      asm volatile("prmt.b32 %0, %1, %2, %3;" : "=r"(v) : "r"(v), "r"(w), "r"(k1));
    if (OP == FFMA)    { float f = __uint_as_float(v); asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(f) : "f"(f), "f"(__uint_as_float(k2)), "f"(__uint_as_float(k1))); v = __float_as_uint(f); }
    if (OP == FMNMX)   { float f; asm volatile("max.f32 %0, %1, %2;" : "=f"(f) : "f"(__uint_as_float(v)), "f"(__uint_as_float(w))); w = v; v = __float_as_uint(f); }
}

constexpr int CH = 4;      // chains per op kind
constexpr int INNER = 32;

template <int A, int B, int NA, int NB>
__global__ void __launch_bounds__(1024) k_pair(unsigned *out, int iters, unsigned one, unsigned seed, long long *cycles)
{
    unsigned va[CH], wa[CH], vb[CH], wb[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) { va[c] = seed + threadIdx.x + c; wa[c] = seed * 3 + c; vb[c] = seed * 7 + threadIdx.x * 3 + c; wb[c] = seed * 5 + c; }
    const unsigned k1 = seed | 0x00010001u, k2 = 0x3C003C00u ^ (one - 1u);
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int s = 0; s < INNER; ++s) {
#pragma unroll
            for (int c = 0; c < CH; ++c) {
#pragma unroll
                for (int r = 0; r < NA; ++r) step<A>(va[c], wa[c], k1, k2, one);
#pragma unroll
                for (int r = 0; r < NB; ++r) step<B>(vb[c], wb[c], k1, k2, one);
            }
        }
    }
    long long t1 = clock64();
    unsigned acc = 0;
#pragma unroll
    for (int c = 0; c < CH; ++c) acc ^= va[c] ^ wa[c] ^ vb[c] ^ wb[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

static bool g_first = true;
template <int A, int B, int NA, int NB>
int run(int nsm, unsigned *d_out, long long *d_cyc, int threads, int iters)
{
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    k_pair<A, B, NA, NB><<<nsm, threads>>>(d_out, iters / 4 + 1, 1u, 12345u, d_cyc);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    k_pair<A, B, NA, NB><<<nsm, threads>>>(d_out, iters, 1u, 12345u, d_cyc);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    static long long h[1024];
    CK(cudaMemcpy(h, d_cyc, sizeof(long long) * nsm, cudaMemcpyDeviceToHost));
    double cyc = 0; for (int b = 0; b < nsm; ++b) cyc += (double)h[b]; cyc /= nsm;
    const double instr = (double)threads * iters * INNER * CH * (NA + NB);
    printf("%s  {\"a\": \"%s\", \"na\": %d, \"b\": \"%s\", \"nb\": %d, \"threads_per_sm\": %d, "
           "\"thread_instr_per_clk_per_sm\": %.2f, \"ms\": %.3f, \"mhz\": %.0f}\n", g_first ? "" : ",",
           names[A], NA, NB ? names[B] : "-", NB, threads, instr / cyc, ms, cyc / (ms * 1e-3) / 1e6);
    g_first = false;
    return 0;
}

#define SOLO(A) rc |= run<A, A, 1, 0>(nsm, d_out, d_cyc, threads, iters);
#define PAIR(A, B) rc |= run<A, B, 1, 1>(nsm, d_out, d_cyc, threads, iters);

int main(int argc, char **argv)
{
    CK(cudaSetDevice(0));
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    const int nsm = prop.multiProcessorCount;
    const int threads = argc > 1 ? atoi(argv[1]) : 1024;
    const int iters = argc > 2 ? atoi(argv[2]) : 2000;
    unsigned *d_out; long long *d_cyc;
    CK(cudaMalloc(&d_out, sizeof(unsigned) * nsm * 1024));
    CK(cudaMalloc(&d_cyc, sizeof(long long) * 1024));
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"results\": [\n", prop.name, nsm);
    int rc = 0;
    SOLO(VIADDMNMX) SOLO(VIMNMX) SOLO(VIMNMX3) SOLO(VIADD16) SOLO(IMAD) SOLO(HFMA2R) SOLO(HADD2) SOLO(VHMNMX)
    SOLO(LOP3) SOLO(IADD3) SOLO(SHF) SOLO(PRMT) SOLO(FFMA) SOLO(FMNMX)
    PAIR(VIADDMNMX, VIMNMX) PAIR(VIADDMNMX, VIMNMX3) PAIR(VIADDMNMX, VIADD16) PAIR(VIADDMNMX, IMAD)
    PAIR(VIADDMNMX, HFMA2R) PAIR(VIADDMNMX, HADD2) PAIR(VIADDMNMX, VHMNMX) PAIR(VIADDMNMX, LOP3)
    PAIR(VIADDMNMX, IADD3) PAIR(VIADDMNMX, FFMA) PAIR(VIADDMNMX, FMNMX)
    PAIR(VIMNMX, IMAD) PAIR(VIMNMX, HFMA2R) PAIR(VIMNMX, HADD2) PAIR(VIMNMX, VHMNMX) PAIR(VIMNMX, FFMA)
    PAIR(VIMNMX, VIADD16) PAIR(VIMNMX, LOP3) PAIR(VIMNMX, IADD3) PAIR(VIMNMX, VIMNMX3)
    PAIR(VIADD16, IMAD) PAIR(VIADD16, HFMA2R) PAIR(VIADD16, FFMA)
    PAIR(IMAD, HFMA2R) PAIR(IMAD, FFMA) PAIR(IMAD, LOP3) PAIR(IMAD, IADD3)
    PAIR(HFMA2R, VHMNMX) PAIR(HFMA2R, FFMA) PAIR(HFMA2R, HADD2) PAIR(HADD2, VHMNMX)
    PAIR(FFMA, FMNMX) PAIR(FFMA, LOP3) PAIR(LOP3, IADD3)
    // register-operand pressure: 3 register sources vs 2 + immediate, alone and next to the FMA-side add
    SOLO(VIADDMNMX_IMM) PAIR(VIADDMNMX_IMM, VIADD16) PAIR(VIADDMNMX_IMM, VIMNMX3) PAIR(VIADDMNMX_IMM, IMAD)
    printf("]}\n");
    return rc;
}
