/*
 * vfgs_b200.h -- additive C-ABI of libvfgs_b200.so (plain pointers and sizes, no C++/torch types).
 *
 * The reference drives its hardware layer one picture line at a time (src/vfgs_main.c:664-682
 * calling vfgs_add_grain_line, src/vfgs_hw.c:288). A GPU wants whole frames, so the frame loop of
 * src/vfgs_main.c:771-790 (read -> vfgs_add_grain -> optional yuv_to_8bit -> write) gets these
 * batch entry points. They share one hardware state (the mirror of src/vfgs_hw.c:49-63) with the
 * drop-in setters of vfgs_hw.h and leave the LFSR registers exactly where the reference's line walk
 * would, so frame calls and line calls can be interleaved bit-exactly.
 *
 * Frame layout ("packed planar", the .yuv file layout of src/yuv.c:162-214): per frame the Y plane
 * (width x height), then U, then V (each (width/subx) x (height/suby)), rows tightly packed, frames
 * back to back. Samples are uint8 when the configured depth is 8 and little-endian uint16 when it
 * is 10. out_depth = 0 keeps the input depth; out_depth = 8 with 10-bit input fuses the
 * (v + 2) >> 2 conversion of src/yuv.c:216-258 into the store.
 *
 * All functions return VFGS_B200_OK or an error code; vfgs_b200_last_error() gives the text.
 * Nothing here falls back to the CPU.
 */
#ifndef VFGS_B200_H
#define VFGS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VFGS_B200_OK         0
#define VFGS_B200_ERR_ARG    1 /* bad argument (size, depth, null pointer, aliasing) */
#define VFGS_B200_ERR_STATE  2 /* hw state the reference would assert on (src/vfgs_hw.c:167-170) */
#define VFGS_B200_ERR_CUDA   3 /* CUDA runtime error, text in vfgs_b200_last_error() */

/* Planes with explicit strides, for callers whose frames are not packed (src/yuv.c:54-87 pads
 * strides to 64 samples). All strides are in BYTES. */
typedef struct vfgs_b200_planes {
	void*   y;
	void*   u;
	void*   v;
	int64_t stride_y;      /* between luma lines */
	int64_t stride_c;      /* between chroma lines */
	int64_t frame_stride;  /* between consecutive frames, same for the three planes */
} vfgs_b200_planes;

/* Bind the library to a CUDA device (default: the current device at first use). */
int vfgs_b200_init(int device);
/* Hardware state back to the power-on values of src/vfgs_hw.c:49-63. */
int vfgs_b200_reset(void);
const char* vfgs_b200_last_error(void);

/* Bytes of one packed planar frame at the given depth with the configured chroma subsampling. */
size_t vfgs_b200_frame_bytes(int width, int height, int depth);

/* nframes packed planar frames already in DEVICE memory; asynchronous on `stream` (a cudaStream_t,
 * NULL = default stream). in == out (every plane identical, same strides) is allowed when the depths
 * match: the kernels then read nothing but what they overwrite themselves; only components with
 * sample-adaptive pattern selection AND 8-sample blocks or ragged / unaligned rows take a detour through a
 * scratch buffer (two extra passes over the data). Input and output that overlap in any other way are
 * refused with VFGS_B200_ERR_ARG. */
int vfgs_b200_add_grain_frames_device(const void* in, void* out, int nframes, int width, int height,
                                      int out_depth, void* stream);

/* Same with explicit planes/strides (device memory). Any sample-aligned pointers and strides are accepted and give the
 * same output; the speed depends on them: planes, line strides and the frame stride that are multiples of 32 bytes
 * (16 for 8-bit samples), with widths that are multiples of 16 samples, take the 16-samples-per-lane kernels
 * (packed frames of the usual picture sizes in cudaMalloc'd memory qualify); multiples of 16 bytes / 8 samples the
 * 8-samples-per-lane form; anything else the EDGE variant. */
int vfgs_b200_add_grain_planes_device(const vfgs_b200_planes* in, const vfgs_b200_planes* out,
                                      int nframes, int width, int height, int out_depth, void* stream);

/* nframes packed planar frames in HOST memory; returns when `out` is complete. Frames are cut into
 * chunks that flow through a ring of device buffers on three streams (H2D / kernels / D2H) so the
 * copies of neighbouring chunks overlap the kernels. Give it pinned buffers (vfgs_b200_host_alloc)
 * for the copies to be asynchronous; pageable buffers work, but the CUDA runtime then stages them
 * and the overlap is lost. in == out is allowed when the depths match. */
int vfgs_b200_add_grain_frames_host(const void* in, void* out, int nframes, int width, int height,
                                    int out_depth);

/* Advance the LFSR registers as if `nframes` frames of this size had been processed (GF(2)
 * jump-ahead, no GPU work): this is how a rank that owns frames [k, k+m) of a sequence gets the
 * state of frame k without touching frames 0..k-1. */
int vfgs_b200_skip_frames(int64_t nframes, int width, int height);

/* Raw LFSR registers in the order rnd, rnd_up, line_rnd, line_rnd_up (src/vfgs_hw.c:52-55). */
void vfgs_b200_get_lfsr(uint32_t regs[4]);
void vfgs_b200_set_lfsr(const uint32_t regs[4]);

/* Copy of the mirrored hardware state, laid out like the reference's statics (src/vfgs_hw.c:49-63):
 * int8 pattern[2][9][64][64], uint8 sLUT[3][256], uint8 pLUT[3][256], uint32 rnd, rnd_up, line_rnd,
 * line_rnd_up, then int scale_shift, bs, Y_min, Y_max, C_min, C_max, csubx, csuby. Returns the size
 * needed; nothing is written when cap is smaller. Needs no GPU. */
size_t vfgs_b200_get_state(void* dst, size_t cap);

/* Page-locked host memory for the host entry point. */
void* vfgs_b200_host_alloc(size_t bytes);
void  vfgs_b200_host_free(void* p);

/* Number of CUDA kernels this library has launched since load (for launch accounting). */
uint64_t vfgs_b200_launch_count(void);

/* Measurement aid: with timing enabled every grain-kernel launch is bracketed by CUDA events on its
 * own stream; vfgs_b200_kernel_time() waits for them and returns the accumulated device time and
 * launch count since timing was (re-)enabled. */
int vfgs_b200_kernel_timing(int enable);
int vfgs_b200_kernel_time(double* total_ms, uint64_t* launches);

/* Test aid: kernel selection. 0 = automatic (fast kernel for single-pattern components -- its EDGE variant
 * for ragged / unaligned rows --, gather kernel for sample-adaptive ones, general kernel for sample-adaptive
 * components on ragged / unaligned rows), 1 = general kernel for everything, 2 = gather kernel wherever it
 * can run. All of them are CUDA paths. */
void vfgs_b200_force_general_kernel(int mode);

/* Frame pipeline behind include/yuv.h: out[0] = frames processed, out[1] = batches flushed. */
void vfgs_b200_pipeline_stats(unsigned long long out[2]);

/* Geometry of the last grain kernel launch: out[0]=grid, out[1]=block, out[2]=dynamic smem bytes,
 * out[3]=SM count of the bound device, out[4]=which grain kernels the last frame call launched
 * (bit 0 fast, bit 1 general, bit 2 gather, bit 3 the fast kernel's EDGE variant). */
void vfgs_b200_last_launch(int out[5]);

#ifdef __cplusplus
}
#endif
#endif /* VFGS_B200_H */
