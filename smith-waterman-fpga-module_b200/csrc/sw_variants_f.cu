/* sw_variants_f.cu -- ahead-of-time instances of the strip kernel (one slice of the variant table). */
#include "sw_variants.h"

namespace swk {
static const VariantEntry g_part[] = {
    // warp-wide systolic groups (intra-task regime) and the small-R latency variants (P ~ query length)
    SW_VARIANT_S16FD(16, 1, 32, 3),
    SW_VARIANT_S16F(8, 2, 32, 3),
    SW_VARIANT_S16D(1, 1, 32, 4),
    SW_VARIANT_S16D(2, 1, 32, 4),
    SW_VARIANT_S16D(4, 1, 32, 4),
    SW_VARIANT_S16D(8, 1, 32, 4),
    SW_VARIANT_S16D(8, 1, 16, 4),
    SW_VARIANT_S16D(16, 1, 8, 4),
    SW_VARIANT_S16D(2, 2, 32, 4),
};
VariantPart sw_variants_part_f() { return {g_part, (int)(sizeof(g_part) / sizeof(g_part[0]))}; }
}  // namespace swk
