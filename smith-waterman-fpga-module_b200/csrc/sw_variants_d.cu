/* sw_variants_d.cu -- ahead-of-time instances of the strip kernel (one slice of the variant table). */
#include "sw_variants.h"

namespace swk {
static const VariantEntry g_part[] = {
    SW_VARIANT_S16F2(25, 3, 1, 2),
    SW_VARIANT_S16F2(38, 2, 1, 2),
    SW_VARIANT_S16(25, 4, 1, 2),
};
VariantPart sw_variants_part_d() { return {g_part, (int)(sizeof(g_part) / sizeof(g_part[0]))}; }
}  // namespace swk
