/*
 * vfgs_hw.h -- drop-in for the reference hardware-layer interface, B200 (sm_100a) back end.
 *
 * The ten entry points below are exactly the ones the reference declares in src/vfgs_hw.h:51-62
 * (same names, argument order, argument meaning and "void + assert" error behaviour), so the
 * unmodified firmware layer (src/vfgs_fw.c) and CLI (src/vfgs_main.c) link against
 * libvfgs_b200.so instead of src/vfgs_hw.c without a source change. Types are spelled with
 * <stdint.h> names; they are ABI-identical to the reference's int8/uint8/uint32 macros
 * (src/vfgs_hw.h:40-47).
 *
 * Behind them the state of src/vfgs_hw.c:49-63 is mirrored on the host and uploaded to the GPU on
 * demand; grain synthesis itself runs only as CUDA kernels (there is no CPU fallback: if no CUDA
 * device is usable the first call that needs one prints the CUDA error and aborts, like a failed
 * assert in the reference).
 *
 * Additive, batch-oriented entry points (whole frames, device-resident or pipelined from host
 * memory) are declared in vfgs_b200.h.
 */
#ifndef VFGS_B200_VFGS_HW_H
#define VFGS_B200_VFGS_HW_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VFGS_MAX_PATTERNS 8 /* reference src/vfgs_hw.h:49 */

/* replaces vfgs_hw.c:314-318 -- P: 64x64 int8, row-major; copied before returning */
void vfgs_set_luma_pattern(int index, int8_t* P);
/* replaces vfgs_hw.c:320-325 -- P: (64/csuby) rows of stride (64/csuby), first 64/csubx bytes of
 * each row are used; repacked with the subsampling in force at the time of the call */
void vfgs_set_chroma_pattern(int index, int8_t* P);
/* replaces vfgs_hw.c:327-331 -- c in 0..2, 256 entries, indexed by the 8 MSBs of the sample */
void vfgs_set_scale_lut(int c, uint8_t lut[]);
/* replaces vfgs_hw.c:333-337 -- pattern index = lut[i] >> 4 */
void vfgs_set_pattern_lut(int c, uint8_t lut[]);

/* replaces vfgs_hw.c:339-344 -- all four LFSR registers <- seed << 1 */
void vfgs_set_seed(uint32_t seed);
/* replaces vfgs_hw.c:346-350 -- 2 <= shift < 8; effective shift is shift + 6 - (depth - 8) */
void vfgs_set_scale_shift(int shift);
/* replaces vfgs_hw.c:352-362 -- 8 or 10; re-bases the effective scale shift */
void vfgs_set_depth(int depth);
/* replaces vfgs_hw.c:364-380 -- 0: clip to [0,255]<<bs, else Y [16,235]<<bs, C [16,240]<<bs */
void vfgs_set_legal_range(int legal);
/* replaces vfgs_hw.c:382-388 -- each 1 or 2 */
void vfgs_set_chroma_subsampling(int subx, int suby);

/* replaces vfgs_hw.c:288-312 -- one picture line of HOST memory, in place. Compatibility path:
 * the line is staged to the GPU, processed by the same kernels as the frame entry points and
 * copied back before the call returns (correct, not fast; use vfgs_b200.h for throughput). */
void vfgs_add_grain_line(void* Y, void* U, void* V, int y, int width);

#ifdef __cplusplus
}
#endif
#endif /* VFGS_B200_VFGS_HW_H */
