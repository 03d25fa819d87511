/*
 * sw_jit.h -- run-time specialisation of the strip kernel for a handle's gap penalties.
 *
 * The reference loads its penalties at run time (ld_penalties, ScoreBank_v2.v:34,161).  The packed
 * DPX instructions take register or immediate operands only, and the immediate form is ~7 % faster
 * (DESIGN.md section 5.1), so for any penalty set that is not compiled in the one kernel variant a
 * job uses is compiled with the penalties as immediates: NVRTC (loaded with dlopen -- the library
 * has no link-time dependency on it) -> cubin -> cudaLibraryLoadData.  Results are cached in memory
 * and on disk ($SW_B200_JIT_CACHE, default ~/.cache/sw_b200).  If NVRTC is missing or the compile
 * fails the caller simply keeps the run-time-operand instance: still the GPU path, never a fallback
 * to anything else.
 */
#ifndef SW_JIT_H_
#define SW_JIT_H_

#include "sw_kernels.h"

/* cudaKernel_t (as void *) of sw_strip_kernel<RS, S, G, ArithS16, false, BT, MINB, goe, ge>, or null.
 * msg (optional) receives a one-line reason when null is returned. */
void *sw_jit_strip_kernel(const SwStripVariant *v, int goe, int ge, char *msg, size_t msg_cap);

/* 1 if NVRTC could be loaded in this process */
int sw_jit_available(void);

#endif
