// pipe_peak.cu -- measures the issue rate of the packed-16-bit DPX / integer / half2
// instructions the Smith-Waterman cell update is built from (SURVEY section 8(d):
// "R_int = measured thread-instructions / clock / SM").  Standalone: nvcc -arch=sm_100a.
// Prints one JSON object.  Not part of the product library.
#include <cstdio>
#include <cstdint>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
    fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)

constexpr int CHAINS = 8;
constexpr int INNER = 64;     // unrolled ops per chain per outer iteration

enum Op { OP_VIADDMAX, OP_VIADDMAX_RELU, OP_VIMAX, OP_VIMAX3, OP_VADD2, OP_IMAD_ADD, OP_IADD,
          OP_HFMA2_RELU, OP_HADD2, OP_HMAX2, OP_MIX_5ALU_1IMAD, OP_MIX_HALF, OP_SWCELL, OP_SWCELL_IMAD,
          OP_SWCELL_LDS, OP_COUNT };
static const char *op_names[] = {"viaddmax_s16x2", "viaddmax_s16x2_relu", "vimax_s16x2", "vimax3_s16x2",
    "vadd2", "imad_add32", "iadd32", "hfma2_relu", "hadd2", "hmax2", "mix_5dpx_1imad", "mix_half_3alu_3fma",
    "sw_cell_6op", "sw_cell_5op_1imad", "sw_cell_6op_lds"};
// thread-instructions counted per "inner step" per chain for each op
static const int op_instr[] = {1, 1, 2, 1, 1, 1, 1, 1, 1, 1, 6, 6, 6, 6, 6};

__device__ __forceinline__ unsigned imad_add(unsigned a, unsigned one, unsigned c) {
    unsigned d; asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(one), "r"(c)); return d;
}
__device__ __forceinline__ unsigned hfma2_relu(unsigned a, unsigned b, unsigned c) {
    unsigned d; asm volatile("fma.rn.relu.f16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d;
}
__device__ __forceinline__ unsigned hadd2u(unsigned a, unsigned b) {
    unsigned d; asm volatile("add.rn.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d;
}
__device__ __forceinline__ unsigned hmax2u(unsigned a, unsigned b) {
    unsigned d; asm volatile("max.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d;
}

template <int OP>
__global__ void __launch_bounds__(256) k_peak(unsigned *out, int iters, unsigned one, unsigned seed,
                                              long long *cycles)
{
    __shared__ unsigned lds_tab[INNER * CHAINS * 16];
    if (OP == OP_SWCELL_LDS) {
        for (int i = threadIdx.x; i < INNER * CHAINS * 16; i += blockDim.x) lds_tab[i] = (i * 2654435761u) & 0x00070007u;
        __syncthreads();
    }
    unsigned v[CHAINS], w[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) { v[c] = seed + threadIdx.x + c; w[c] = seed * 3 + c; }
    const unsigned k1 = seed | 0x00010001u, k2 = (seed >> 3) | 0x00020002u;
    const unsigned h_one = 0x3C003C00u;     // half2(1,1)
    unsigned best = 0;
    const unsigned zero = one - 1u;   // opaque 0: a literal 0 makes ptxas emit a PRMT per use
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int s = 0; s < INNER; ++s) {
#pragma unroll
            for (int c = 0; c < CHAINS; ++c) {
                if (OP == OP_VIADDMAX)        v[c] = __viaddmax_s16x2(v[c], k1, k2);
                if (OP == OP_VIADDMAX_RELU)   v[c] = __viaddmax_s16x2_relu(v[c], k1, k2);
                if (OP == OP_VIMAX)         { v[c] = __vmaxs2(v[c], w[c]); w[c] = __vmins2(w[c], v[(c + 1) % CHAINS]); }
                if (OP == OP_VIMAX3)          v[c] = __vimax3_s16x2(v[c], k1, w[c]);
                if (OP == OP_VADD2)           v[c] = __vadd2(v[c], k1);
                if (OP == OP_IMAD_ADD)        v[c] = imad_add(v[c], one, k1);
                if (OP == OP_IADD)            v[c] = v[c] + (k1 ^ v[(c + 1) % CHAINS]);
                if (OP == OP_HFMA2_RELU)      v[c] = hfma2_relu(v[c], h_one, k1);
                if (OP == OP_HADD2)           v[c] = hadd2u(v[c], k1);
                if (OP == OP_HMAX2)           v[c] = hmax2u(v[c], w[c]);
                if (OP == OP_MIX_5ALU_1IMAD) {
                    unsigned m = __viaddmax_s16x2_relu(v[c], k1, zero);
                    unsigned i_ = __vmaxs2(w[c], k2);
                    unsigned j = imad_add(i_, one, k1);
                    w[c] = __viaddmax_s16x2(m, k2, j);
                    v[c] = __vmaxs2(m, i_);
                    best = __vmaxs2(best, v[c]);
                }
                if (OP == OP_MIX_HALF) {
                    unsigned m = hfma2_relu(v[c], h_one, k1);        // fma pipe
                    unsigned i_ = __vmaxs2(w[c], k2);                // alu
                    unsigned j = hadd2u(i_, k1);                     // fma pipe
                    unsigned mo = hadd2u(m, k2);                     // fma pipe
                    w[c] = __vmaxs2(mo, j);                          // alu
                    v[c] = __vmaxs2(m, i_);                          // alu
                    best = __vmaxs2(best, v[c]);                     // alu
                }
                if (OP == OP_SWCELL || OP == OP_SWCELL_IMAD || OP == OP_SWCELL_LDS) {
                    // one column step of a register strip: v = H (diag for next row), w = G (left gap value)
                    unsigned sc = k1;
                    if (OP == OP_SWCELL_LDS) sc = lds_tab[((s * CHAINS + c) << 4) + ((threadIdx.x + it) & 15)];
                    unsigned hd = v[(c + CHAINS - 1) % CHAINS];          // H of the row above, previous column
                    unsigned gu = w[(c + CHAINS - 1) % CHAINS];
                    unsigned m = __viaddmax_s16x2_relu(hd, sc, zero);
                    unsigned i_ = __vmaxs2(w[c], gu);
                    unsigned j = (OP == OP_SWCELL_IMAD) ? imad_add(i_, one, k2) : __vadd2(i_, k2);
                    w[c] = __viaddmax_s16x2(m, k1, j);
                    v[c] = __vmaxs2(m, i_);
                    best = __vmaxs2(best, v[c]);
                }
            }
        }
    }
    long long t1 = clock64();
    unsigned acc = best;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) acc ^= v[c] ^ w[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int OP>
int run(int nsm, unsigned *d_out, long long *d_cyc, int blocks_per_sm, int iters, bool first)
{
    const int threads = 256;
    const int blocks = nsm * blocks_per_sm;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    k_peak<OP><<<blocks, threads>>>(d_out, iters / 8 + 1, 1u, 12345u, d_cyc);   // warm-up
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    k_peak<OP><<<blocks, threads>>>(d_out, iters, 1u, 12345u, d_cyc);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms = 0; CK(cudaEventElapsedTime(&ms, e0, e1));
    static long long h_cyc[4096];
    CK(cudaMemcpy(h_cyc, d_cyc, sizeof(long long) * blocks, cudaMemcpyDeviceToHost));
    double cyc = 0; for (int b = 0; b < blocks; ++b) cyc += (double)h_cyc[b]; cyc /= blocks;
    const double instr_per_block = (double)threads * iters * INNER * CHAINS * op_instr[OP];
    // every SM runs blocks_per_sm blocks concurrently for ~cyc cycles
    const double per_clk_sm = instr_per_block * blocks_per_sm / cyc;
    const double ginstr_s = instr_per_block * blocks / (ms * 1e-3) / 1e9;
    printf("%s  {\"op\": \"%s\", \"thread_instr_per_clk_per_sm\": %.2f, \"tera_thread_instr_per_s\": %.3f, "
           "\"ms\": %.3f, \"avg_block_cycles\": %.0f, \"implied_mhz\": %.0f}\n", first ? "" : ",",
           op_names[OP], per_clk_sm, ginstr_s / 1e3, ms, cyc, cyc / (ms * 1e-3) / 1e6);
    return 0;
}

int main(int argc, char **argv)
{
    int dev = 0; CK(cudaSetDevice(dev));
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, dev));
    const int nsm = prop.multiProcessorCount;
    const int bps = (argc > 1) ? atoi(argv[1]) : 4;        // 4 x 256 threads = 32 warps / SM
    const int iters = (argc > 2) ? atoi(argv[2]) : 400;
    unsigned *d_out; long long *d_cyc;
    CK(cudaMalloc(&d_out, sizeof(unsigned) * nsm * bps * 256));
    CK(cudaMalloc(&d_cyc, sizeof(long long) * 4096));
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"blocks_per_sm\": %d, \"threads_per_block\": 256, \"results\": [\n",
           prop.name, nsm, bps);
    int rc = 0;
    rc |= run<OP_VIADDMAX>(nsm, d_out, d_cyc, bps, iters, true);
    rc |= run<OP_VIADDMAX_RELU>(nsm, d_out, d_cyc, bps, iters, false);
    rc |= run<OP_VIMAX>(nsm, d_out, d_cyc, bps, iters, false);
    rc |= run<OP_VIMAX3>(nsm, d_out, d_cyc, bps, iters, false);
    rc |= run<OP_VADD2>(nsm, d_out, d_cyc, bps, iters, false);
    rc |= run<OP_IMAD_ADD>(nsm, d_out, d_cyc, bps, iters, false);
    rc |= run<OP_IADD>(nsm, d_out, d_cyc, bps, iters, false);
    rc |= run<OP_HFMA2_RELU>(nsm, d_out, d_cyc, bps, iters, false);
    rc |= run<OP_HADD2>(nsm, d_out, d_cyc, bps, iters, false);
    rc |= run<OP_HMAX2>(nsm, d_out, d_cyc, bps, iters, false);
    rc |= run<OP_MIX_5ALU_1IMAD>(nsm, d_out, d_cyc, bps, iters / 4, false);
    rc |= run<OP_MIX_HALF>(nsm, d_out, d_cyc, bps, iters / 4, false);
    rc |= run<OP_SWCELL>(nsm, d_out, d_cyc, bps, iters / 4, false);
    rc |= run<OP_SWCELL_IMAD>(nsm, d_out, d_cyc, bps, iters / 4, false);
    rc |= run<OP_SWCELL_LDS>(nsm, d_out, d_cyc, bps, iters / 4, false);
    printf("]}\n");
    return rc;
}
