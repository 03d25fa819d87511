/* sw_variants.h -- table entry of one strip-kernel variant; the instances live in sw_variants_*.cu
 * (split over several translation units so that ptxas runs in parallel). */
#ifndef SW_VARIANTS_H_
#define SW_VARIANTS_H_

#include "sw_kernels.h"
#include "sw_strip.cuh"

namespace swk {

constexpr int kBT = 128;

typedef void (*StripFn)(const StripArgs);

struct VariantEntry {
    SwStripVariant info;
    StripFn fn;        // exact arithmetic, run-time penalties
    StripFn fn_w12;    // W-bit wrap-then-clamp
    StripFn fn_fixed;  // exact arithmetic, gap penalties kFixedGoe / kFixedGe as immediates (or null)
    StripFn fn_fixed2; // same for the second compiled-in set kFixed2Goe / kFixed2Ge (or null)
    StripFn fn_direct; // DIRECT: codes formed on the fly from the 2-bit records (or null)
};

// the reference's default gap penalties: gap_open -12, gap_extend -4  =>  goe = -16, ge = -4
constexpr int kFixedGoe = -16, kFixedGe = -4;
// second compiled-in set: gap_open -8, gap_extend -4, the parameters of the reference's swalign
// golden vectors (data/sw_testing.txt: first gap residue costs -12)
constexpr int kFixed2Goe = -12, kFixed2Ge = -4;

#define SW_K(RS, S, G, W12, MINB, ...) sw_strip_kernel<RS, S, G, ArithS16, W12, kBT, MINB, ##__VA_ARGS__>
#define SW_INFO(RS, S, G, MINB, D) {RS * S, G, kBT, S, MINB, "strip_s16x2_R" #RS "x" #S "_G" #G, D, 4, 0}
// run-time penalties + W-bit
#define SW_VARIANT_S16(RS, S, G, MINB) \
    { SW_INFO(RS, S, G, MINB, 0), SW_K(RS, S, G, false, MINB), SW_K(RS, S, G, true, MINB), nullptr, nullptr, nullptr }
// + an instance with the default gap penalties as immediates
#define SW_VARIANT_S16F(RS, S, G, MINB) \
    { SW_INFO(RS, S, G, MINB, 0), SW_K(RS, S, G, false, MINB), SW_K(RS, S, G, true, MINB), \
      SW_K(RS, S, G, false, MINB, kFixedGoe, kFixedGe), nullptr, nullptr }
// + instances for both compiled-in gap penalty sets
#define SW_VARIANT_S16F2(RS, S, G, MINB) \
    { SW_INFO(RS, S, G, MINB, 0), SW_K(RS, S, G, false, MINB), SW_K(RS, S, G, true, MINB), \
      SW_K(RS, S, G, false, MINB, kFixedGoe, kFixedGe), SW_K(RS, S, G, false, MINB, kFixed2Goe, kFixed2Ge), nullptr }
// run-time penalties + W-bit + DIRECT (small-batch / latency path)
#define SW_VARIANT_S16D(RS, S, G, MINB) \
    { SW_INFO(RS, S, G, MINB, 1), SW_K(RS, S, G, false, MINB), SW_K(RS, S, G, true, MINB), nullptr, nullptr, \
      SW_K(RS, S, G, false, MINB, 0, 0, true) }
#define SW_VARIANT_S16FD(RS, S, G, MINB) \
    { SW_INFO(RS, S, G, MINB, 1), SW_K(RS, S, G, false, MINB), SW_K(RS, S, G, true, MINB), \
      SW_K(RS, S, G, false, MINB, kFixedGoe, kFixedGe), nullptr, SW_K(RS, S, G, false, MINB, 0, 0, true) }

// experimental: U columns per trip of the step loop (the register fix-up moves at the loop's back
// edge are amortised over more columns); selectable by name for A/B measurements
#define SW_VARIANT_S16F_U(RS, S, G, MINB, U) \
    { {RS * S, G, kBT, S, MINB, "strip_s16x2_R" #RS "x" #S "_G" #G "_U" #U, 0, U, 0}, \
      SW_K(RS, S, G, false, MINB, 0, 0, false, U), SW_K(RS, S, G, true, MINB, 0, 0, false, U), \
      SW_K(RS, S, G, false, MINB, kFixedGoe, kFixedGe, false, U), nullptr, nullptr }

// interior-trip flags FL (sw_strip.cuh, SW_FAST_LOOP) on top of U
#define SW_VARIANT_S16F_UF(RS, S, G, MINB, U, FL) \
    { {RS * S, G, kBT, S, MINB, "strip_s16x2_R" #RS "x" #S "_G" #G "_U" #U "_F" #FL, 0, U, FL}, \
      SW_K(RS, S, G, false, MINB, 0, 0, false, U, FL), SW_K(RS, S, G, true, MINB, 0, 0, false, U, FL), \
      SW_K(RS, S, G, false, MINB, kFixedGoe, kFixedGe, false, U, FL), nullptr, nullptr }

struct VariantPart { const VariantEntry *v; int n; };
VariantPart sw_variants_part_a();
VariantPart sw_variants_part_b();
VariantPart sw_variants_part_c();
VariantPart sw_variants_part_d();
VariantPart sw_variants_part_e();
VariantPart sw_variants_part_f();
VariantPart sw_variants_part_g();
VariantPart sw_variants_part_h();

}  // namespace swk
#endif
