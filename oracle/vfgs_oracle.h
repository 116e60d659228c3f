/*
 * vfgs_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement ("oracle") of the VFGS hardware-layer hot path, reference src/vfgs_hw.c:74-388
 * plus the output-depth conversion src/yuv.c:216-258. Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may use it, and only as the checker. The product
 * (versatilefilmgrain_b200/) never includes, links or calls anything in oracle/.
 *
 * Parity status: PINNED BY EXECUTION. The reference ships no golden vectors,
 * so this restatement is checked against the reference itself, compiled unmodified into
 * oracle/_ref/libvfgs_ref.so (tests/test_oracle_vs_reference.py, every cfg/ file), and against the
 * fixtures that library generated (tests/golden/, made by tests/golden/make_golden.py).
 */
#ifndef VFGS_ORACLE_H
#define VFGS_ORACLE_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct vfgs_oracle vfgs_oracle;

vfgs_oracle* oracle_new(void);               /* power-on state of vfgs_hw.c:49-63 */
void oracle_free(vfgs_oracle* o);

/* The nine configuration entry points of vfgs_hw.h:51-60, same argument meaning. They return 0,
 * or -1 where the reference would trip an assert (hw.c:316,322,329,335,348,354,384-385). */
int oracle_set_luma_pattern(vfgs_oracle* o, int index, const int8_t* P);
int oracle_set_chroma_pattern(vfgs_oracle* o, int index, const int8_t* P);
int oracle_set_scale_lut(vfgs_oracle* o, int c, const uint8_t lut[256]);
int oracle_set_pattern_lut(vfgs_oracle* o, int c, const uint8_t lut[256]);
int oracle_set_seed(vfgs_oracle* o, uint32_t seed);
int oracle_set_scale_shift(vfgs_oracle* o, int shift);
int oracle_set_depth(vfgs_oracle* o, int depth);
int oracle_set_legal_range(vfgs_oracle* o, int legal);
int oracle_set_chroma_subsampling(vfgs_oracle* o, int subx, int suby);

/* vfgs_add_grain_line (hw.c:288-312): one picture line, in place, serial register bookkeeping. */
int oracle_add_grain_line(vfgs_oracle* o, void* Y, void* U, void* V, int y, int width);

/* Whole frames in closed ("parallel") form: every block's LFSR state comes from a GF(2) jump-ahead
 * of t = (f*(R-1) + r)*nb + b steps from the register value at entry, nothing is carried from
 * sample to sample. Equivalent to driving oracle_add_grain_line over y = 0..height-1 of each frame
 * the way vfgs_main.c:664-682 does, followed by yuv_to_8bit when out_depth == 8 < input depth.
 * Planes are tightly packed (stride == width), Y then U then V, frames back to back; in == out is
 * allowed when the depths match. Advances the LFSR registers exactly like the line walk would. */
int oracle_add_grain_frames(vfgs_oracle* o, const void* in, void* out, int nframes, int width,
                            int height, int out_depth);

/* LFSR helpers: one step (hw.c:74-79), n steps by matrix powers, raw register access. */
uint32_t oracle_lfsr_step(uint32_t x);
uint32_t oracle_lfsr_jump(uint32_t x, uint64_t n);
void oracle_get_lfsr(const vfgs_oracle* o, uint32_t regs[4]); /* rnd, rnd_up, line_rnd, line_rnd_up */
void oracle_set_lfsr(vfgs_oracle* o, const uint32_t regs[4]);

/* Offset decode (hw.c:99-138): out = {sign, ox, oy}. */
void oracle_block_offsets(int c, uint32_t state, int subx, int suby, int out[3]);

/* State dump in the layout of refh_state (oracle/ref_harness.c) for direct comparison. */
size_t oracle_state_size(void);
void oracle_get_state(const vfgs_oracle* o, void* dst);

#ifdef __cplusplus
}
#endif
#endif
