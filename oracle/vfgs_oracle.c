/*
 * vfgs_oracle.c -- TEST INFRASTRUCTURE ONLY (see vfgs_oracle.h). CPU restatement of the VFGS
 * hardware-layer hot path; every function names the reference lines it restates.
 * Written from the algorithm description in SURVEY.md section 8(a), not from the reference text:
 * grain is evaluated per sample in closed form (block state -> offsets -> LUTs -> pattern ->
 * vertical blend -> edge filter -> scale/add/clip) instead of through the reference's two-block
 * shift register.
 */
#include "vfgs_oracle.h"
#include <stdlib.h>
#include <string.h>

#define NSLOT 9 /* 8 settable patterns + the always-zero slot 8 (hw.c:49) */

struct vfgs_oracle {
	int8_t  pat[2][NSLOT][64][64];
	uint8_t slut[3][256];
	uint8_t plut[3][256];
	uint32_t rnd, rnd_up, line_rnd, line_rnd_up;
	int scale_shift, bs;
	int y_min, y_max, c_min, c_max;
	int csubx, csuby;
};

/* ---------------------------------------------------------------- LFSR (hw.c:70-79) */

/* 31-bit Fibonacci LFSR kept in bits 31..1; bit 0 only trails. New top bit = b1 ^ b29. */
uint32_t oracle_lfsr_step(uint32_t x)
{
	uint32_t fb = ((x >> 1) ^ (x >> 29)) & 1u;
	return (x >> 1) | (fb << 31);
}

/* GF(2) transition matrix in row form: bit i of M*x = parity(row[i] & x). */
typedef struct { uint32_t row[32]; } gf2m;

static uint32_t gf2_apply(const gf2m* m, uint32_t x)
{
	uint32_t y = 0;
	for (int i = 0; i < 32; i++)
		y |= (uint32_t)(__builtin_popcount(m->row[i] & x) & 1) << i;
	return y;
}

static void gf2_mul(gf2m* out, const gf2m* a, const gf2m* b) /* out = a*b (apply b first) */
{
	gf2m t;
	for (int i = 0; i < 32; i++) {
		uint32_t acc = 0, sel = a->row[i];
		for (int j = 0; j < 32; j++)
			if ((sel >> j) & 1u) acc ^= b->row[j];
		t.row[i] = acc;
	}
	*out = t;
}

static gf2m g_pow2[64]; /* g_pow2[k] = M^(2^k) */
static int g_pow2_ready = 0;

static void lfsr_tables(void)
{
	if (g_pow2_ready) return;
	for (int i = 0; i < 31; i++) g_pow2[0].row[i] = 1u << (i + 1);
	g_pow2[0].row[31] = (1u << 1) | (1u << 29);
	for (int k = 1; k < 64; k++) gf2_mul(&g_pow2[k], &g_pow2[k - 1], &g_pow2[k - 1]);
	g_pow2_ready = 1;
}

uint32_t oracle_lfsr_jump(uint32_t x, uint64_t n)
{
	lfsr_tables();
	for (int k = 0; n; k++, n >>= 1)
		if (n & 1u) x = gf2_apply(&g_pow2[k], x);
	return x;
}

/* ---------------------------------------------------------------- offsets (hw.c:99-138) */

static int bin13(uint32_t f10) { return (int)((f10 * 13u) >> 10); } /* 0..12 */
static int bin12(uint32_t f10) { return (int)((f10 * 12u) >> 10); } /* 0..11 */

void oracle_block_offsets(int c, uint32_t v, int subx, int suby, int out[3])
{
	uint32_t sign_bit, fx, fy;
	int stepx, stepy;
	if (c == 0) {            /* hw.c:103-109 */
		sign_bit = v >> 31;
		fx = v & 0x3ff;
		fy = (v >> 14) & 0x3ff;
		stepx = stepy = 4;
	} else if (c == 1) {     /* hw.c:118-124: the y field wraps around the word */
		sign_bit = v >> 2;
		fx = (v >> 10) & 0x3ff;
		fy = ((v >> 24) & 0xff) | ((v & 3u) << 8);
		stepx = 4 / subx; stepy = 4 / suby;
	} else {                 /* hw.c:131-137 */
		sign_bit = v >> 15;
		fx = (v >> 20) & 0x3ff;
		fy = (v >> 4) & 0x3ff;
		stepx = 4 / subx; stepy = 4 / suby;
	}
	out[0] = (sign_bit & 1u) ? -1 : 1;
	out[1] = bin13(fx) * stepx;
	out[2] = bin12(fy) * stepy;
}

/* ---------------------------------------------------------------- state + setters (hw.c:314-388) */

vfgs_oracle* oracle_new(void)
{
	vfgs_oracle* o = (vfgs_oracle*)calloc(1, sizeof(*o));
	if (!o) return NULL;
	o->rnd = o->rnd_up = o->line_rnd = o->line_rnd_up = 0xdeadbeefu; /* hw.c:52-55 */
	o->scale_shift = 11;                                             /* hw.c:56 */
	o->bs = 0;
	o->y_min = o->c_min = 0; o->y_max = o->c_max = 255;
	o->csubx = o->csuby = 2;
	return o;
}

void oracle_free(vfgs_oracle* o) { free(o); }

int oracle_set_luma_pattern(vfgs_oracle* o, int index, const int8_t* P) /* hw.c:314-318 */
{
	if (index < 0 || index >= 8) return -1;
	memcpy(o->pat[0][index], P, 64 * 64);
	return 0;
}

int oracle_set_chroma_pattern(vfgs_oracle* o, int index, const int8_t* P) /* hw.c:320-325 */
{
	if (index < 0 || index >= 8) return -1;
	int rows = 64 / o->csuby, src_stride = 64 / o->csuby, ncopy = 64 / o->csubx;
	for (int r = 0; r < rows; r++)
		memcpy(o->pat[1][index][r], P + (size_t)src_stride * r, (size_t)ncopy);
	return 0;
}

int oracle_set_scale_lut(vfgs_oracle* o, int c, const uint8_t lut[256]) /* hw.c:327-331 */
{
	if (c < 0 || c > 2) return -1;
	memcpy(o->slut[c], lut, 256);
	return 0;
}

int oracle_set_pattern_lut(vfgs_oracle* o, int c, const uint8_t lut[256]) /* hw.c:333-337 */
{
	if (c < 0 || c > 2) return -1;
	for (int i = 0; i < 256; i++)
		if ((lut[i] >> 4) >= NSLOT) return -1; /* would index past pattern[..][9] in hw.c:218 */
	memcpy(o->plut[c], lut, 256);
	return 0;
}

int oracle_set_seed(vfgs_oracle* o, uint32_t seed) /* hw.c:339-344 */
{
	o->rnd = o->rnd_up = o->line_rnd = o->line_rnd_up = seed << 1;
	return 0;
}

int oracle_set_scale_shift(vfgs_oracle* o, int shift) /* hw.c:346-350 */
{
	if (shift < 2 || shift >= 8) return -1;
	o->scale_shift = shift + 6 - o->bs;
	return 0;
}

int oracle_set_depth(vfgs_oracle* o, int depth) /* hw.c:352-362: shift follows the depth change */
{
	if (depth != 8 && depth != 10) return -1;
	int nbs = depth - 8;
	o->scale_shift = (o->scale_shift - nbs + o->bs) & 0xff; /* uint8 in the reference */
	o->bs = nbs;
	return 0;
}

int oracle_set_legal_range(vfgs_oracle* o, int legal) /* hw.c:364-380 */
{
	o->y_min = o->c_min = legal ? 16 : 0;
	o->y_max = legal ? 235 : 255;
	o->c_max = legal ? 240 : 255;
	return 0;
}

int oracle_set_chroma_subsampling(vfgs_oracle* o, int subx, int suby) /* hw.c:382-388 */
{
	if ((subx != 1 && subx != 2) || (suby != 1 && suby != 2)) return -1;
	o->csubx = subx; o->csuby = suby;
	return 0;
}

void oracle_get_lfsr(const vfgs_oracle* o, uint32_t r[4])
{
	r[0] = o->rnd; r[1] = o->rnd_up; r[2] = o->line_rnd; r[3] = o->line_rnd_up;
}

void oracle_set_lfsr(vfgs_oracle* o, const uint32_t r[4])
{
	o->rnd = r[0]; o->rnd_up = r[1]; o->line_rnd = r[2]; o->line_rnd_up = r[3];
}

/* dump in refh_state layout (oracle/ref_harness.c) */
typedef struct {
	int8_t  pattern[2][NSLOT][64][64];
	uint8_t slut[3][256];
	uint8_t plut[3][256];
	uint32_t rnd, rnd_up, line_rnd, line_rnd_up;
	int scale_shift, bs, y_min, y_max, c_min, c_max, csubx, csuby;
} oracle_dump;

size_t oracle_state_size(void) { return sizeof(oracle_dump); }

void oracle_get_state(const vfgs_oracle* o, void* dst)
{
	oracle_dump* d = (oracle_dump*)dst;
	memcpy(d->pattern, o->pat, sizeof(o->pat));
	memcpy(d->slut, o->slut, sizeof(o->slut));
	memcpy(d->plut, o->plut, sizeof(o->plut));
	d->rnd = o->rnd; d->rnd_up = o->rnd_up; d->line_rnd = o->line_rnd; d->line_rnd_up = o->line_rnd_up;
	d->scale_shift = o->scale_shift; d->bs = o->bs;
	d->y_min = o->y_min; d->y_max = o->y_max; d->c_min = o->c_min; d->c_max = o->c_max;
	d->csubx = o->csubx; d->csuby = o->csuby;
}

/* ---------------------------------------------------------------- one component line */

#define MAXW 16384
#define MAXNB (MAXW / 16 + 2)

static int floor_shift(int a, int s) { return (a + (1 << (s - 1))) >> s; } /* hw.c:43 */

/*
 * One line of one component (restates add_grain_block, hw.c:140-284, sample by sample).
 *   src/dst  line of this component (may alias); src_depth in {8,10}; dst8 != 0 stores
 *            (v+2)>>2 as bytes (yuv.c:216-258), else same width as the source
 *   y        LUMA line number (drives j = y&15 and the overlap rule), cw = in-picture samples
 *   st_cur/st_up  LFSR state of block 0 for the current / upper block-row; block b uses b steps on
 * Samples right of the picture inside the last (partial) block are taken as 0 (the reference reads
 * whatever sits in the stride padding there; only a one-sample-wide last block can see it).
 */
static void component_line(const vfgs_oracle* o, int c, const void* src, void* dst, int dst8,
                           int y, int cw, int nb, uint32_t st_cur, uint32_t st_up)
{
	static __thread int16_t G[MAXW + 32];
	static __thread int16_t Gf[MAXW + 32];
	static __thread uint8_t it[MAXW + 32];
	const int subx = c ? o->csubx : 1, suby = c ? o->csuby : 1;
	const int n = 16 / subx;                 /* samples per block, hw.c:209 */
	const int j = y & 15;
	const int bs = o->bs, ss = o->scale_shift;
	const int lo = (c ? o->c_min : o->y_min) << bs, hi = (c ? o->c_max : o->y_max) << bs;
	const int8_t (*bank)[64][64] = o->pat[c ? 1 : 0];
	const uint8_t* s8 = (const uint8_t*)src;
	const uint16_t* s16 = (const uint16_t*)src;

	/* overlap weights, hw.c:173-188 */
	int w_cur = 0, w_up = 0;
	if (y > 15 && j == 0) { w_cur = suby > 1 ? 20 : 12; w_up = suby > 1 ? 20 : 24; }
	else if (y > 15 && j == 1) { w_cur = 24; w_up = 12; }

	/* intensities, hw.c:211 */
	for (int k = 0; k < nb * n; k++) {
		int v = k < cw ? (bs ? s16[k] : s8[k]) : 0;
		it[k] = (uint8_t)(bs ? v >> bs : v);
	}

	/* raw grain per sample, hw.c:190-236 */
	for (int b = 0; b < nb; b++) {
		int oc[3], ou[3];
		oracle_block_offsets(c, st_cur, subx, suby, oc);
		oracle_block_offsets(c, st_up, subx, suby, ou);
		const int row_cur = oc[2] + j / suby, row_up = ou[2] + (16 + j) / suby;
		for (int i = 0; i < n; i++) {
			int k = b * n + i;
			int pi = o->plut[c][it[k]] >> 4;
			int p = bank[pi][row_cur][oc[1] + i] * oc[0];
			if (w_cur)
				p = floor_shift(p * w_cur + bank[pi][row_up][ou[1] + i] * w_up * ou[0], 5);
			G[k] = (int16_t)p;
		}
		st_cur = oracle_lfsr_step(st_cur);   /* hw.c:309-310 */
		st_up = oracle_lfsr_step(st_up);
	}

	/* block-edge filter, hw.c:250-259: every boundary except x = 0, taps read unfiltered grain */
	memcpy(Gf, G, sizeof(int16_t) * (size_t)(nb * n));
	G[nb * n] = 0;
	for (int b = 1; b < nb; b++) {
		int r = b * n, l = r - 1;
		Gf[l] = (int16_t)floor_shift(G[l - 1] + 3 * G[l] + G[r], 2);
		Gf[r] = (int16_t)floor_shift(G[l] + 3 * G[r] + G[r + 1], 2);
	}

	/* scale, add, clip, store, hw.c:260-268 (+ yuv.c:231 for the 8-bit output) */
	for (int k = 0; k < cw; k++) {
		int v = bs ? s16[k] : s8[k];
		int g = floor_shift(o->slut[c][it[k]] * Gf[k], ss);
		int r = v + g;
		r = r > hi ? hi : r;
		r = r < lo ? lo : r;
		if (dst8 && bs)      ((uint8_t*)dst)[k] = (uint8_t)((r + 2) >> 2);
		else if (bs)         ((uint16_t*)dst)[k] = (uint16_t)r;
		else                 ((uint8_t*)dst)[k] = (uint8_t)r;
	}
}

static int state_ok(const vfgs_oracle* o, int width)
{
	/* hw.c:168-170 */
	if (width <= 128 || width > MAXW) return 0;
	if (o->bs != 0 && o->bs != 2) return 0;
	if (o->scale_shift + o->bs < 8 || o->scale_shift + o->bs > 13) return 0;
	return 1;
}

/* ---------------------------------------------------------------- line entry (hw.c:288-312) */

int oracle_add_grain_line(vfgs_oracle* o, void* Y, void* U, void* V, int y, int width)
{
	if (!state_ok(o, width)) return -1;
	const int nb = (width + 15) / 16;
	if (y && (y & 15) == 0) { o->line_rnd_up = o->line_rnd; o->line_rnd = o->rnd; }
	component_line(o, 0, Y, Y, 0, y, width, nb, o->line_rnd, o->line_rnd_up);
	if (!((y & 1) && o->csuby > 1)) { /* hw.c:164-165 */
		int cw = width / o->csubx;
		component_line(o, 1, U, U, 0, y, cw, nb, o->line_rnd, o->line_rnd_up);
		component_line(o, 2, V, V, 0, y, cw, nb, o->line_rnd, o->line_rnd_up);
	}
	o->rnd = oracle_lfsr_jump(o->line_rnd, (uint64_t)nb);
	o->rnd_up = oracle_lfsr_jump(o->line_rnd_up, (uint64_t)nb);
	return 0;
}

/* ---------------------------------------------------------------- frames, closed form */

int oracle_add_grain_frames(vfgs_oracle* o, const void* in, void* out, int nframes, int width,
                            int height, int out_depth)
{
	if (!state_ok(o, width) || height < 1) return -1;
	const int in_depth = 8 + o->bs;
	if (out_depth == 0) out_depth = in_depth;
	if (out_depth != in_depth && !(out_depth == 8 && in_depth == 10)) return -1;
	const int dst8 = out_depth == 8 && in_depth == 10;
	if (dst8 && in == out) return -1;

	const int nb = (width + 15) / 16, R = (height + 15) / 16;
	const int cw = width / o->csubx, ch = height / o->csuby;   /* yuv.c:72-77 */
	const size_t isz = in_depth > 8 ? 2 : 1, osz = out_depth > 8 ? 2 : 1;
	const size_t ysam = (size_t)width * height, csam = (size_t)cw * ch;
	const uint32_t s0 = o->line_rnd;                            /* epoch state */

	for (int f = 0; f < nframes; f++) {
		const uint8_t* fi = (const uint8_t*)in + (size_t)f * (ysam + 2 * csam) * isz;
		uint8_t* fo = (uint8_t*)out + (size_t)f * (ysam + 2 * csam) * osz;
		for (int y = 0; y < height; y++) {
			const int r = y >> 4;
			/* block 0 of block-row r of frame f sits t0 steps after the epoch state; the first
			 * row of a frame re-uses the last row's state of the frame before (hw.c:291-298) */
			const uint64_t t0 = ((uint64_t)f * (uint64_t)(R - 1) + (uint64_t)r) * (uint64_t)nb;
			const uint32_t st_cur = oracle_lfsr_jump(s0, t0);
			const uint32_t st_up = r ? oracle_lfsr_jump(s0, t0 - (uint64_t)nb) : st_cur; /* unused when y<16 */
			component_line(o, 0, fi + (size_t)y * width * isz, fo + (size_t)y * width * osz, dst8,
			               y, width, nb, st_cur, st_up);
			if ((y % o->csuby) == 0 && y / o->csuby < ch) {
				const size_t cl = (size_t)(y / o->csuby) * cw;
				for (int c = 1; c <= 2; c++) {
					const size_t base = ysam + (size_t)(c - 1) * csam + cl;
					component_line(o, c, fi + base * isz, fo + base * osz, dst8, y, cw, nb, st_cur, st_up);
				}
			}
		}
	}

	/* leave the registers where the reference's line walk would (hw.c:291-298, 309-310) */
	if (nframes > 0) {
		const uint64_t adv = (uint64_t)nframes * (uint64_t)(R - 1) * (uint64_t)nb;
		if (R >= 2) {
			o->line_rnd_up = oracle_lfsr_jump(s0, adv - (uint64_t)nb);
			o->line_rnd = oracle_lfsr_jump(s0, adv);
		}
		o->rnd = oracle_lfsr_jump(o->line_rnd, (uint64_t)nb);
		o->rnd_up = oracle_lfsr_jump(o->line_rnd_up, (uint64_t)nb);
	}
	return 0;
}
