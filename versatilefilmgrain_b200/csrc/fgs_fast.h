// fgs_fast.h -- the throughput path of the grain kernel: components whose pattern LUT selects one
// single pattern slot (every AFGS1 config, most FGC-SEI configs), rows
// aligned for 128-bit access, widths a multiple of 8 samples.
//
// Same decomposition as fgs_task.h (warp-task = 256 samples x the lines of one stripe, lane = 8
// samples, neighbours' edge samples recomputed from their own LFSR window), but the per-sample work
// is cut to the bone:
//   * with one pattern slot the unscaled grain does not depend on the sample, so a lane's 8 grain
//     bytes per line are one contiguous octet of a pattern row: two/three 32-bit shared loads;
//   * the block's random sign (vfgs_hw.c:218 "* s") is folded into WHICH COPY of the pattern is
//     read: the table image holds +pattern and -pattern (the host only takes this path when no
//     pattern byte is -128), so no per-sample sign multiply is left;
//   * the scale LUT (vfgs_hw.c:239) is replicated per lane in shared memory ([256][32] words,
//     word = sLUT[Y] | sLUT[U] << 8 | sLUT[V] << 16): a lane only ever touches its own bank, so the
//     256-entry lookup with random intensities is conflict-free by construction;
//   * add + clip (vfgs_hw.c:265-267) run on two samples at a time with the packed 16-bit min/max
//     instructions (VIADDMNMX / VIMNMX .S16x2); 10-bit samples stay packed in their load words.
// Host-compilable like fgs_task.h (tests/emu) -- the helpers below emulate the few PTX instructions.
#pragma once
#include "fgs_task.h"

namespace vfgs {

// ---- byte permute and packed 16-bit helpers ------------------------------------------------
// PTX prmt.b32, default mode: selector nibble n (bits 2:0) picks byte n of {b,a}; nibble bit 3
// replicates that byte's sign bit instead.
VFGS_HD uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel)
{
#if defined(__CUDA_ARCH__)
	uint32_t d;
	asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
	return d;
#else
	const uint64_t src = ((uint64_t)b << 32) | a;
	uint32_t d = 0;
	for (int i = 0; i < 4; i++) {
		const uint32_t n = (sel >> (4 * i)) & 0xf;
		uint32_t byte = (uint32_t)(src >> (8 * (n & 7))) & 0xff;
		if (n & 8) byte = (byte & 0x80) ? 0xff : 0x00;
		d |= byte << (8 * i);
	}
	return d;
#endif
}
// sign-extended byte e (0..7) of the octet {w1,w0}
template <int E>
VFGS_HD int octet_byte(uint32_t w0, uint32_t w1)
{
	constexpr uint32_t n = (uint32_t)E, s = 8u | (uint32_t)E;
	return (int)prmt(w0, w1, n | (s << 4) | (s << 8) | (s << 12));
}
VFGS_HD uint32_t add_max_s16x2(uint32_t a, uint32_t b, uint32_t c) // per half: max(a + b, c), signed
{
#if defined(__CUDA_ARCH__)
	return __viaddmax_s16x2(a, b, c);
#else
	uint32_t r = 0;
	for (int i = 0; i < 2; i++) {
		int16_t x = (int16_t)((int16_t)(a >> (16 * i)) + (int16_t)(b >> (16 * i)));
		const int16_t y = (int16_t)(c >> (16 * i));
		x = x > y ? x : y;
		r |= (uint32_t)(uint16_t)x << (16 * i);
	}
	return r;
#endif
}
VFGS_HD uint32_t min_s16x2(uint32_t a, uint32_t b)
{
#if defined(__CUDA_ARCH__)
	return __vmins2(a, b);
#else
	uint32_t r = 0;
	for (int i = 0; i < 2; i++) {
		const int16_t x = (int16_t)(a >> (16 * i)), y = (int16_t)(b >> (16 * i));
		r |= (uint32_t)(uint16_t)(x < y ? x : y) << (16 * i);
	}
	return r;
#endif
}
VFGS_HD uint32_t min_u16x2(uint32_t a, uint32_t b)
{
#if defined(__CUDA_ARCH__)
	return __vminu2(a, b);
#else
	uint32_t r = 0;
	for (int i = 0; i < 2; i++) {
		const uint16_t x = (uint16_t)(a >> (16 * i)), y = (uint16_t)(b >> (16 * i));
		r |= (uint32_t)(x < y ? x : y) << (16 * i);
	}
	return r;
#endif
}

// ---- shared-memory image of the fast path --------------------------------------------------
// [0, kLutBytes)            uint32 lut[256][32]   per-lane replicated scale LUT (built by the CTA)
// [kLutBytes, +fblob_bytes) copy of the global image: uint32 compact_lut[256], then for each component
//                           its pattern slot twice (+ and -), rows packed to fpat_stride bytes
constexpr int kLutBytes = 256 * 32 * 4;

VFGS_HD uint32_t lds32(const uint8_t* smem, int off) { return *(const uint32_t*)(smem + off); }
VFGS_HD int lds_s8(const uint8_t* smem, int off) { return (int)(int8_t)smem[off]; }
VFGS_HD int lds_u8(const uint8_t* smem, int off) { return (int)smem[off]; }

// 8 consecutive pattern bytes at byte offset `off`; ALIGNED: off % 4 == 0 (luma-type components),
// else off % 4 in {0, 2} (horizontally subsampled chroma: ox is a multiple of 2).
template <bool ALIGNED>
VFGS_HD void octet(const uint8_t* smem, int off, uint32_t& w0, uint32_t& w1)
{
	if (ALIGNED) {
		w0 = lds32(smem, off); w1 = lds32(smem, off + 4);
	} else {
		const int a = off & ~3, sh = (off & 3) * 8;
		const uint32_t x0 = lds32(smem, a), x1 = lds32(smem, a + 4), x2 = lds32(smem, a + 8);
#if defined(__CUDA_ARCH__)
		w0 = __funnelshift_r(x0, x1, sh); w1 = __funnelshift_r(x1, x2, sh);
#else
		w0 = sh ? (x0 >> sh) | (x1 << (32 - sh)) : x0;
		w1 = sh ? (x1 >> sh) | (x2 << (32 - sh)) : x1;
#endif
	}
}

// Per-lane constants of a warp-task (pattern byte offsets have the block's sign folded in).
struct FastLane {
	int own;                 // the lane's octet, pattern row of line j = 0 of the current block
	int lh, rh;              // halo bytes: last column of block b-1 / first column of block b+1
	int stride;              // pattern row pitch
	int lut;                 // byte offset of this lane's LUT column + component byte
	bool has_left, has_right;
	int ss, rnd;
	uint32_t lo2, hi2;       // clip range replicated in both 16-bit halves (8-bit input: plain ints)
};
// Same offsets for the block-row above, only alive while the overlap lines are processed.
struct FastUp {
	int own, lh, rh;
};

template <int E>
VFGS_HD int blend(int cur, uint32_t u0, uint32_t u1, int w_cur, int w_up) // vfgs_hw.c:225
{
	return (cur * w_cur + octet_byte<E>(u0, u1) * w_up + 16) >> 5;
}

// One line of one lane. raw: the lane's 8 samples as loaded (IN16: 4 words of two 10-bit samples,
// else 2 words of four bytes). outw: result words ready to store (16-bit out: 4 words, 8-bit: 2).
// rc: byte offset of this line's row inside the current block's window; w_cur != 0 selects the
// vertical-overlap blend with row offset ru of the upper block's window.
template <bool IN16, bool OUT8, int NSH>
VFGS_HD void fast_line(const FastLane& L, const uint8_t* smem, int rc, int w_cur, int w_up, const FastUp& U, int ru,
                       const uint32_t raw[4], uint32_t outw[4])
{
	constexpr bool ALIGNED = NSH == 4;
	uint32_t c0, c1;
	octet<ALIGNED>(smem, L.own + rc, c0, c1);
	int g[8];
	g[0] = octet_byte<0>(c0, c1); g[1] = octet_byte<1>(c0, c1); g[2] = octet_byte<2>(c0, c1); g[3] = octet_byte<3>(c0, c1);
	g[4] = octet_byte<4>(c0, c1); g[5] = octet_byte<5>(c0, c1); g[6] = octet_byte<6>(c0, c1); g[7] = octet_byte<7>(c0, c1);
	int hl = L.has_left ? lds_s8(smem, L.lh + rc) : 0;
	int hr = L.has_right ? lds_s8(smem, L.rh + rc) : 0;

	// vertical overlap with the block-row above (vfgs_hw.c:173-188, 223-229)
	if (w_cur) {
		uint32_t u0, u1;
		octet<ALIGNED>(smem, U.own + ru, u0, u1);
		g[0] = blend<0>(g[0], u0, u1, w_cur, w_up); g[1] = blend<1>(g[1], u0, u1, w_cur, w_up);
		g[2] = blend<2>(g[2], u0, u1, w_cur, w_up); g[3] = blend<3>(g[3], u0, u1, w_cur, w_up);
		g[4] = blend<4>(g[4], u0, u1, w_cur, w_up); g[5] = blend<5>(g[5], u0, u1, w_cur, w_up);
		g[6] = blend<6>(g[6], u0, u1, w_cur, w_up); g[7] = blend<7>(g[7], u0, u1, w_cur, w_up);
		if (L.has_left) hl = (hl * w_cur + lds_s8(smem, U.lh + ru) * w_up + 16) >> 5;
		if (L.has_right) hr = (hr * w_cur + lds_s8(smem, U.rh + ru) * w_up + 16) >> 5;
	}

	// block-edge filter (vfgs_hw.c:250-259), taps read unfiltered grain
	const int f0 = (hl + 3 * g[0] + g[1] + 2) >> 2;
	const int f7 = (g[6] + 3 * g[7] + hr + 2) >> 2;
	g[0] = L.has_left ? f0 : g[0];
	g[7] = L.has_right ? f7 : g[7];

	if (IN16) {
		uint32_t r[4];
#pragma unroll
		for (int k = 0; k < 4; k++) {
			// LUT index = (sample >> 2) & 0xff (vfgs_hw.c:211), times the 128-byte LUT row pitch
			const int s_lo = lds_u8(smem, L.lut + (int)((raw[k] << 5) & 0x7f80u));
			const int s_hi = lds_u8(smem, L.lut + (int)((raw[k] >> 11) & 0x7f80u));
			const int d_lo = (s_lo * g[2 * k] + L.rnd) >> L.ss;      // vfgs_hw.c:263
			const int d_hi = (s_hi * g[2 * k + 1] + L.rnd) >> L.ss;
			const uint32_t d2 = prmt((uint32_t)d_lo, (uint32_t)d_hi, 0x5410);
			// samples above 0x3fff clip to the ceiling whatever the grain: cap them so the signed 16-bit add cannot wrap
			const uint32_t v2 = min_u16x2(raw[k], 0x3fff3fffu);
			r[k] = min_s16x2(add_max_s16x2(v2, d2, L.lo2), L.hi2);  // vfgs_hw.c:265
		}
		if (OUT8) {
#pragma unroll
			for (int k = 0; k < 4; k++) r[k] = ((r[k] + 0x00020002u) >> 2) & 0x00ff00ffu; // yuv.c:231
			outw[0] = prmt(r[0], r[1], 0x6420);
			outw[1] = prmt(r[2], r[3], 0x6420);
		} else {
			outw[0] = r[0]; outw[1] = r[1]; outw[2] = r[2]; outw[3] = r[3];
		}
	} else {
		const int lo = (int)(L.lo2 & 0xffff), hi = (int)(L.hi2 & 0xffff);
		int o[8];
#pragma unroll
		for (int e = 0; e < 8; e++) {
			const int v = (int)((raw[e >> 2] >> ((e & 3) * 8)) & 0xff);
			const int s = lds_u8(smem, L.lut + (v << 7));
			int x = v + ((s * g[e] + L.rnd) >> L.ss);
			x = x > hi ? hi : x;
			o[e] = x < lo ? lo : x;
		}
		outw[0] = (uint32_t)o[0] | ((uint32_t)o[1] << 8) | ((uint32_t)o[2] << 16) | ((uint32_t)o[3] << 24);
		outw[1] = (uint32_t)o[4] | ((uint32_t)o[5] << 8) | ((uint32_t)o[6] << 16) | ((uint32_t)o[7] << 24);
	}
}

// Pattern byte offset of (block window, column) inside the shared image.
VFGS_HD int window_off(const FgsParams& p, int c, const BlockOfs& o, int col)
{
	return p.fpat_off[c][o.sign < 0 ? 1 : 0] + o.oy * p.fpat_stride[c] + o.ox + col;
}

constexpr int kFastLB = 4; // lines whose loads are issued back to back

// kFastLB consecutive lines of the component (nl of them exist). ovl: this is the first batch of a
// stripe that has a block-row above it: line 0 (and line 1 when the component is not vertically
// subsampled) blend with the upper block's window U.
template <bool IN16, bool OUT8, int NSH>
VFGS_HD void fast_batch(const FastLane& L, const FastUp& U, const uint8_t* smem, int rc0, int ysh, bool ovl,
                        const uint8_t* src, uint8_t* dst, long long in_pitch, long long out_pitch, int nl)
{
	constexpr int OB = (IN16 && !OUT8) ? 2 : 1;
	uint32_t raw[kFastLB][4];
#pragma unroll
	for (int q = 0; q < kFastLB; q++) {
		// a short last batch re-reads its last line instead of branching around the load
		const int qq = q < nl ? q : nl - 1;
		if (IN16) ld_global_16(src + qq * in_pitch, raw[q]);
		else ld_global_8(src + qq * in_pitch, raw[q]);
	}
#pragma unroll
	for (int q = 0; q < kFastLB; q++) {
		uint32_t w[4];
		int w_cur = 0, w_up = 0, ru = 0;
		if (q == 0 && ovl) { w_cur = ysh ? 20 : 12; w_up = ysh ? 20 : 24; ru = (16 >> ysh) * L.stride; }
		if (q == 1 && ovl && !ysh) { w_cur = 24; w_up = 12; ru = 17 * L.stride; }
		fast_line<IN16, OUT8, NSH>(L, smem, rc0 + q * L.stride, w_cur, w_up, U, ru, raw[q], w);
		if (q < nl) {
			if (OB == 2) st_global_16(dst + q * out_pitch, w);
			else st_global_8(dst + q * out_pitch, w);
		}
	}
}

template <bool IN16, bool OUT8, int NSH>
VFGS_HD void fast_task_body(const FgsParams& p, const uint8_t* smem, const TaskGeom& t, int lane)
{
	const int c = t.c;
	const Plane& pl = p.comp[c];
	const int ysh = (c && p.suby > 1) ? 1 : 0;
	constexpr int n = 1 << NSH;
	const int k0 = t.seg * kSegSamples + lane * kSamplesPerLane;
	if (k0 >= pl.width) return;

	// component lines of this stripe (whole stripes only: the host sends partial line ranges to
	// the general kernel)
	const int cl0 = (t.r * 16) >> ysh;
	int cl1 = cl0 + (16 >> ysh);
	if (cl1 > pl.lines) cl1 = pl.lines;
	if (cl0 >= cl1) return;

	const int b = k0 >> NSH;
	const int i0 = k0 & (n - 1);
	FastLane L;
	L.has_left = (i0 == 0) && (b > 0);
	L.has_right = (i0 + kSamplesPerLane == n) && (b + 1 < p.nb);
	L.stride = p.fpat_stride[c];
	L.lut = lane * 4 + c;
	L.ss = p.ss; L.rnd = 1 << (p.ss - 1);
	L.lo2 = (uint32_t)p.lo[c] * 0x00010001u; L.hi2 = (uint32_t)p.hi[c] * 0x00010001u;

	const int srow = t.r - p.stream_row0;
	const uint32_t* row_cur = p.streams + ((long long)t.f * p.stream_rows + srow) * p.wpr;
	L.own = window_off(p, c, decode_offsets(c, stream_window(row_cur, b), p.subx, p.suby), i0);
	L.lh = L.rh = 0;
	if (L.has_left) L.lh = window_off(p, c, decode_offsets(c, stream_window(row_cur, b - 1), p.subx, p.suby), n - 1);
	if (L.has_right) L.rh = window_off(p, c, decode_offsets(c, stream_window(row_cur, b + 1), p.subx, p.suby), 0);

	constexpr int IB = IN16 ? 2 : 1, OB = (IN16 && !OUT8) ? 2 : 1;
	const long long in_pitch = pl.in_row_bytes, out_pitch = pl.out_row_bytes;
	const uint8_t* src = pl.in + (long long)t.f * p.in_frame_bytes + (long long)cl0 * in_pitch + (long long)k0 * IB;
	uint8_t* dst = pl.out + (long long)t.f * p.out_frame_bytes + (long long)cl0 * out_pitch + (long long)k0 * OB;
	int nl = cl1 - cl0;

	// the first batch of a stripe overlaps the block-row above (never in the first stripe, y <= 15)
	FastUp U;
	U.own = U.lh = U.rh = 0;
	bool ovl = t.r > 0;
	if (ovl) {
		const uint32_t* row_up = row_cur - p.wpr;
		U.own = window_off(p, c, decode_offsets(c, stream_window(row_up, b), p.subx, p.suby), i0);
		if (L.has_left) U.lh = window_off(p, c, decode_offsets(c, stream_window(row_up, b - 1), p.subx, p.suby), n - 1);
		if (L.has_right) U.rh = window_off(p, c, decode_offsets(c, stream_window(row_up, b + 1), p.subx, p.suby), 0);
	}
	int rc = 0;
#pragma unroll 1
	for (; nl > 0; nl -= kFastLB) {
		fast_batch<IN16, OUT8, NSH>(L, U, smem, rc, ysh, ovl, src, dst, in_pitch, out_pitch, nl);
		src += kFastLB * in_pitch; dst += kFastLB * out_pitch;
		rc += kFastLB * L.stride;
		ovl = false;
	}
}

// Dispatch on the component's block size (16 samples: luma and non-subsampled chroma; 8: chroma
// subsampled horizontally).
template <bool IN16, bool OUT8>
VFGS_HD void process_task_fast(const FgsParams& p, const uint8_t* smem, long long task, int lane)
{
	const TaskGeom t = decode_task(p, task);
	if (t.c && p.subx > 1) fast_task_body<IN16, OUT8, 3>(p, smem, t, lane);
	else fast_task_body<IN16, OUT8, 4>(p, smem, t, lane);
}

} // namespace vfgs
