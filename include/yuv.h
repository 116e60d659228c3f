/*
 * yuv.h -- batched, pinned-memory replacement for the reference's planar YUV frame I/O
 * (reference src/yuv.h:47-66, src/yuv.c), exported by libvfgs_b200.so.
 *
 * Same seven entry points and the same `yuv` struct layout, so the UNMODIFIED reference CLI
 * (src/vfgs_main.c, which includes its own yuv.h) links against this implementation instead of
 * src/yuv.c. Behind the interface the frame loop of src/vfgs_main.c:771-790
 *
 *     yuv_read -> vfgs_add_grain (one vfgs_add_grain_line per line) -> [yuv_to_8bit] -> yuv_write
 *
 * is turned into a batch pipeline without touching the caller:
 *   - yuv_alloc hands out a ring of page-locked frame slots; yuv_read re-points frame->Y/U/V at the
 *     next slot and reads the whole packed frame with one fread;
 *   - vfgs_add_grain_line on a ring slot only records that the frame wants grain (line 0) -- nothing
 *     is computed per line; yuv_to_8bit records the 8-bit output request;
 *   - yuv_write queues the frame; when the ring is full, when any vfgs_set_* call is about to change
 *     the hardware state (a cfg scheduled mid-stream, src/vfgs_main.c:773-781), or in yuv_free, the
 *     queued frames go through vfgs_b200_add_grain_frames_host (H2D / kernels / D2H on three
 *     streams) and are written to their files in order.
 * Output files are byte-identical to the reference CLI's.
 */
#ifndef VFGS_B200_YUV_H
#define VFGS_B200_YUV_H

#include <stdio.h>

#ifdef __cplusplus
extern "C" {
#endif

#define YUV_420 0
#define YUV_422 1
#define YUV_444 2

/* layout of reference src/yuv.h:47-58 */
typedef struct yuv_s {
	void* Y;
	void* U;
	void* V;
	unsigned short width;
	unsigned short height;
	unsigned short stride;
	unsigned short cwidth;
	unsigned short cheight;
	unsigned short cstride;
	unsigned       depth;
} yuv;

int  yuv_alloc(int width, int height, int depth, int format, yuv* frame); /* replaces yuv.c:54-87 */
void yuv_free(yuv* frame);                                               /* replaces yuv.c:89-95; drains the pipeline */
void yuv_pad(yuv* frame);                                                /* replaces yuv.c:152-160 (unused by the CLI) */
int  yuv_skip(yuv* frame, int n, FILE* file);                            /* replaces yuv.c:97-106 */
int  yuv_read(yuv* frame, FILE* file);                                   /* replaces yuv.c:180-187 */
int  yuv_write(yuv* frame, FILE* file);                                  /* replaces yuv.c:207-214 (deferred, ordered) */
void yuv_to_8bit(yuv* dst, const yuv* src);                              /* replaces yuv.c:216-258 (fused into the kernel store) */

#ifdef __cplusplus
}
#endif
#endif /* VFGS_B200_YUV_H */
