// fgs_gather.h -- grain kernel task code for components whose pattern LUT selects SEVERAL pattern
// slots (sample-adaptive pattern selection: the reference's built-in default SEI with 8 luma
// patterns, cfg/fgs_sei_ff_test5-7 chroma). Same decomposition and the same memory pipeline as the
// single-pattern path (fgs_fast.h: lane = 8 samples, 4 lines in flight with rotating refill, packed
// 16-bit clip), but the grain byte of every sample is a true gather:
//     entry = lut[intensity]            one conflict-free 32-bit shared load; the component has its own
//                                       per-lane replicated table, entry = scale | slot byte offset << 8
//     grain = pattern[slot offset + window row + column]      one byte load, bank conflicts as they fall
// The block's random sign is applied to the fetched byte (FMA pipe), so no negated pattern copies are
// needed and any int8 pattern value is allowed. The neighbour sample an edge filter needs is
// recomputed from the neighbouring block's register and THAT sample's intensity, which sits in the
// adjacent lane's registers: one warp shuffle per side and line (only lanes 0 and 31 read their
// neighbour sample from global memory, which is why this path cannot run in place).
// Restates vfgs_hw.c:140-284 per sample like fgs_task.h; host-compilable for tests/emu.
#pragma once
#include "fgs_fast.h"

namespace vfgs {

// Shared memory: [0, 32 KB * ngather) private LUTs of the gather components (each on a 32 KB
// boundary), then the general table image's pattern slots (gpat_off, relative to the image copy).
#ifndef VFGS_GATHER_LB
#define VFGS_GATHER_LB 2 // 2 lines in flight with 28 warps per SM measured ahead of 3, 4 and 5 with 24 (the kernel is issue-bound;
                         // fewer staging registers leave it free of spills)
#endif
constexpr int kGatherLB = VFGS_GATHER_LB; // lines in flight per lane
static_assert(kGatherLB >= 2, "both vertical-overlap lines of a block-row must fall into the first group of lines");

struct GatherLane {
	smem_addr_t own;        // bank + window row 0 + ox + i0 of the current block (slot offset added per sample)
	smem_addr_t lh, rh;     // neighbour windows: last column of block b-1 / first column of block b+1
	smem_addr_t lut;        // this lane's column of the component's private LUT
	int stride;
	bool has_left, has_right;
	int s_own, s_l, s_r;    // block signs
	int pow16;
	uint32_t lo2, hi2;
};
struct GatherUp {
	smem_addr_t own, lh, rh;
	int s_own, s_l, s_r;
};

VFGS_HD void ld_cached_16(const uint8_t* p, uint32_t r[4])
{
#if defined(__CUDA_ARCH__)
	asm volatile("ld.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "l"(p));
#else
	memcpy(r, p, 16);
#endif
}
VFGS_HD void ld_cached_8(const uint8_t* p, uint32_t r[2])
{
#if defined(__CUDA_ARCH__)
	asm volatile("ld.global.v2.u32 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "l"(p));
#else
	memcpy(r, p, 8);
#endif
}
VFGS_HD void ld_cached_16_if(const uint8_t* p, uint32_t r[4], bool pred)
{
#if defined(__CUDA_ARCH__)
	asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %5, 0;\n\t@q ld.global.v4.u32 {%0,%1,%2,%3}, [%4];\n\t}"
	             : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]) : "l"(p), "r"((uint32_t)pred));
#else
	if (pred) memcpy(r, p, 16);
#endif
}
VFGS_HD void ld_cached_8_if(const uint8_t* p, uint32_t r[2], bool pred)
{
#if defined(__CUDA_ARCH__)
	asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %3, 0;\n\t@q ld.global.v2.u32 {%0,%1}, [%2];\n\t}"
	             : "+r"(r[0]), "+r"(r[1]) : "l"(p), "r"((uint32_t)pred));
#else
	if (pred) memcpy(r, p, 8);
#endif
}
// one sample (IB bytes wide) when pred is set, else 0; volatile so that it is issued where it is
// written (a batch ahead of its use) instead of being sunk next to the use
template <int IB>
VFGS_HD uint32_t ld_sample_if(const uint8_t* p, bool pred)
{
#if defined(__CUDA_ARCH__)
	uint32_t v = 0;
	if (IB == 2)
		asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %2, 0;\n\t@q ld.global.u16 %0, [%1];\n\t}" : "+r"(v) : "l"(p), "r"((uint32_t)pred));
	else
		asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %2, 0;\n\t@q ld.global.u8 %0, [%1];\n\t}" : "+r"(v) : "l"(p), "r"((uint32_t)pred));
	return v;
#else
	if (!pred) return 0;
	return IB == 2 ? (uint32_t)*(const uint16_t*)p : (uint32_t)*p;
#endif
}

#if defined(__CUDA_ARCH__)
constexpr bool kHaloFromMemory = false;
#else
constexpr bool kHaloFromMemory = true; // host build (tests/emu) runs lane by lane: no shuffles
#endif

// Edge samples of the two neighbouring lanes for this line. Device: the left neighbour's last and the
// right neighbour's first sample come out of their registers by shuffle; lane 0 / lane 31 use the value
// they loaded from memory (gl / gr). Samples right of the picture read as 0.
template <bool IN16>
VFGS_HD void neighbour_samples(const uint32_t raw[4], int lane, uint32_t gl, uint32_t gr, bool right_in_picture,
                               uint32_t& vl, uint32_t& vr)
{
#if defined(__CUDA_ARCH__)
	const uint32_t first = IN16 ? raw[0] & 0xffffu : raw[0] & 0xffu;
	const uint32_t last = IN16 ? raw[3] >> 16 : raw[1] >> 24;
	const uint32_t up = __shfl_up_sync(0xffffffffu, last, 1);
	const uint32_t down = __shfl_down_sync(0xffffffffu, first, 1);
	vl = lane == 0 ? gl : up;
	vr = !right_in_picture ? 0u : lane == 31 ? gr : down;
#else
	(void)raw; (void)lane; (void)right_in_picture;
	vl = gl; vr = gr;
#endif
}
// 16-sample blocks: a lane is the first or the second half of a block and has one block edge, so one
// neighbour sample serves. First halves (even lanes) take the last sample of the lane to their left, second
// halves the first sample of the lane to their right: one shuffle, every lane sending what its receiver needs.
template <bool IN16>
VFGS_HD uint32_t neighbour_sample(const uint32_t raw[4], int lane, uint32_t gmem, bool second_half, bool right_in_picture)
{
#if defined(__CUDA_ARCH__)
	const uint32_t first = IN16 ? raw[0] & 0xffffu : raw[0] & 0xffu;
	const uint32_t last = IN16 ? raw[3] >> 16 : raw[1] >> 24;
	const uint32_t got = __shfl_sync(0xffffffffu, second_half ? last : first, second_half ? lane + 1 : lane - 1);
	if (second_half) return !right_in_picture ? 0u : lane == 31 ? gmem : got;
	return lane == 0 ? gmem : got;
#else
	(void)raw; (void)lane;
	return (second_half && !right_in_picture) ? 0u : gmem;
#endif
}

// LUT index bits (intensity * 128) of sample e of a lane's raw words.
template <bool IN16, int E>
VFGS_HD uint32_t index_bits(const uint32_t raw[4])
{
	if (IN16) {
		const uint32_t w = raw[E >> 1];
		return (E & 1) ? (mulhi_u32(w, 1u << 21) & 0x7f80u) : ((w << 5) & 0x7f80u); // ((v >> 2) & 0xff) << 7
	} else {
		const uint32_t w = raw[E >> 2];
		constexpr int sh = (E & 3) * 8;
		return sh >= 7 ? (w >> (sh - 7)) & 0x7f80u : (w << (7 - sh)) & 0x7f80u;      // v << 7
	}
}

// entry -> scale and grain byte of one sample (column E of the lane's window); the block's sign is not applied
template <int E>
VFGS_HD void gather_sample(const GatherLane& L, uint32_t ibits, int rc, int& scale, int& grain)
{
	const uint32_t ent = lds32(L.lut | (smem_addr_t)ibits);
	scale = (int)(ent & 0xff);
	grain = lds_s8(L.own + rc + (smem_addr_t)(ent >> 8) + E);
}
// the same sample's byte in the window of the block above, blended in (vfgs_hw.c:223-229); the weights carry
// the two block signs, so the result is fully signed
template <int E>
VFGS_HD int gather_blend(const GatherLane& L, const GatherUp& U, uint32_t ibits, int ru, int wc_s, int wu_s, int g)
{
	const smem_addr_t off = (smem_addr_t)(lds32(L.lut | (smem_addr_t)ibits) >> 8) + E; // looked up again: overlap lines only
	return (g * wc_s + lds_s8(U.own + ru + off) * wu_s + 16) >> 5;
}

// MERGE (16-sample blocks): the lane's one neighbour sample is vl, its window L.lh / U.lh and sign L.s_l / U.s_l.
// On lines without vertical overlap g[] stays free of the block's sign until the scale multiply (the sign rides
// on the 2^(16 - shift) factor); only the edge filter, whose neighbour tap carries another block's sign, applies
// it explicitly.
template <bool IN16, bool OUT8, bool MERGE>
VFGS_HD void gather_line(const GatherLane& L, const GatherUp& U, int rc, int w_cur, int w_up, int ru, int in_shift,
                         const uint32_t raw[4], uint32_t vl, uint32_t vr, uint32_t outw[4])
{
	int g[8], sc[8];
	gather_sample<0>(L, index_bits<IN16, 0>(raw), rc, sc[0], g[0]);
	gather_sample<1>(L, index_bits<IN16, 1>(raw), rc, sc[1], g[1]);
	gather_sample<2>(L, index_bits<IN16, 2>(raw), rc, sc[2], g[2]);
	gather_sample<3>(L, index_bits<IN16, 3>(raw), rc, sc[3], g[3]);
	gather_sample<4>(L, index_bits<IN16, 4>(raw), rc, sc[4], g[4]);
	gather_sample<5>(L, index_bits<IN16, 5>(raw), rc, sc[5], g[5]);
	gather_sample<6>(L, index_bits<IN16, 6>(raw), rc, sc[6], g[6]);
	gather_sample<7>(L, index_bits<IN16, 7>(raw), rc, sc[7], g[7]);

	// neighbours' edge samples (their own intensity selects their pattern slot); harmless addresses when there
	// is no neighbour (lh/rh fall back to the lane's own window), selected at the end
	const smem_addr_t offl = (smem_addr_t)(lds32(L.lut | (smem_addr_t)(((vl >> in_shift) & 0xffu) << 7)) >> 8);
	int hl = lds_s8(L.lh + rc + offl), hr = 0;
	smem_addr_t offr = 0;
	if (!MERGE) {
		offr = (smem_addr_t)(lds32(L.lut | (smem_addr_t)(((vr >> in_shift) & 0xffu) << 7)) >> 8);
		hr = lds_s8(L.rh + rc + offr);
	}

	int own_sign = L.s_own, mul = L.pow16 * L.s_own; // sign of g[], factor of the scale multiply
	if (w_cur) { // vertical overlap with the block-row above: warp-uniform branch, 2 (1) of 16 (8) lines
		const int wc_s = w_cur * L.s_own, wu_s = w_up * U.s_own;
		g[0] = gather_blend<0>(L, U, index_bits<IN16, 0>(raw), ru, wc_s, wu_s, g[0]);
		g[1] = gather_blend<1>(L, U, index_bits<IN16, 1>(raw), ru, wc_s, wu_s, g[1]);
		g[2] = gather_blend<2>(L, U, index_bits<IN16, 2>(raw), ru, wc_s, wu_s, g[2]);
		g[3] = gather_blend<3>(L, U, index_bits<IN16, 3>(raw), ru, wc_s, wu_s, g[3]);
		g[4] = gather_blend<4>(L, U, index_bits<IN16, 4>(raw), ru, wc_s, wu_s, g[4]);
		g[5] = gather_blend<5>(L, U, index_bits<IN16, 5>(raw), ru, wc_s, wu_s, g[5]);
		g[6] = gather_blend<6>(L, U, index_bits<IN16, 6>(raw), ru, wc_s, wu_s, g[6]);
		g[7] = gather_blend<7>(L, U, index_bits<IN16, 7>(raw), ru, wc_s, wu_s, g[7]);
		hl = (hl * (w_cur * L.s_l) + lds_s8(U.lh + ru + offl) * (w_up * U.s_l) + 16) >> 5;
		if (!MERGE) hr = (hr * (w_cur * L.s_r) + lds_s8(U.rh + ru + offr) * (w_up * U.s_r) + 16) >> 5;
		own_sign = 1; mul = L.pow16;
	} else {
		hl *= L.s_l;
		if (!MERGE) hr *= L.s_r;
	}

	// block-edge filter (vfgs_hw.c:250-259): taps read unfiltered grain; the result goes back into g[]'s sign convention
	if (MERGE) {
		const int a = L.has_right ? g[7] : g[0], b = L.has_right ? g[6] : g[1];
		const int f = ((hl + 2 + own_sign * (3 * a + b)) >> 2) * own_sign;
		g[0] = L.has_left ? f : g[0];
		g[7] = L.has_right ? f : g[7];
	} else {
		const int f0 = ((hl + 2 + own_sign * (3 * g[0] + g[1])) >> 2) * own_sign;
		const int f7 = ((hr + 2 + own_sign * (g[6] + 3 * g[7])) >> 2) * own_sign;
		g[0] = L.has_left ? f0 : g[0];
		g[7] = L.has_right ? f7 : g[7];
	}

	if (IN16) {
		constexpr int kRound = OUT8 ? 0x28000 : 0x8000; // 8-bit output: the + 2 of yuv.c:231 rides on the rounding and the clip range (fast_line)
		uint32_t r[4];
#pragma unroll
		for (int k = 0; k < 4; k++) {
			const int a_lo = sc[2 * k] * (g[2 * k] * mul) + kRound;
			const int a_hi = sc[2 * k + 1] * (g[2 * k + 1] * mul) + kRound;
			const uint32_t d2 = prmt((uint32_t)a_lo, (uint32_t)a_hi, 0x7632);
			const uint32_t v2 = min_u16x2(raw[k], 0x3fff3fffu);
			r[k] = min_s16x2(add_max_s16x2(v2, d2, L.lo2), L.hi2);
		}
		if (OUT8) {
#pragma unroll
			for (int k = 0; k < 4; k++) r[k] >>= 2;
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
			int x = v + ((sc[e] * (g[e] * mul) + 0x8000) >> 16);
			x = x > hi ? hi : x;
			o[e] = x < lo ? lo : x;
		}
		outw[0] = (uint32_t)o[0] | ((uint32_t)o[1] << 8) | ((uint32_t)o[2] << 16) | ((uint32_t)o[3] << 24);
		outw[1] = (uint32_t)o[4] | ((uint32_t)o[5] << 8) | ((uint32_t)o[6] << 16) | ((uint32_t)o[7] << 24);
	}
}

// Window address (without slot offset) and sign of a block from its precomputed table entry
// (FgsParams::woffs, gather format: oy * stride + ox, bit 15 = negative sign).
VFGS_HD smem_addr_t gather_window(smem_addr_t bank, uint32_t entry, int col, int& sign)
{
	sign = (entry & 0x8000u) ? -1 : 1;
	return bank + (smem_addr_t)((entry & 0x7fffu) + (uint32_t)col);
}

template <bool IN16, bool OUT8, int NSH>
VFGS_HD void gather_task_body(const FgsParams& p, smem_addr_t luts, smem_addr_t img, const TaskGeom& t, int lane)
{
	constexpr bool MERGE = NSH == 4; // one block edge per lane
	const int c = t.c;
	const Plane& pl = p.comp[c];
	const int ysh = (c && p.suby > 1) ? 1 : 0;
	constexpr int n = 1 << NSH;
	const int k0 = t.seg * kSegSamples + lane * kSamplesPerLane;
	// lanes right of the picture stay in the loop (the neighbour exchange is a warp shuffle) but
	// neither load nor store
	const bool active = k0 < pl.width;

	const int cl0 = (t.r * 16) >> ysh;
	int cl1 = cl0 + (16 >> ysh);
	if (cl1 > pl.lines) cl1 = pl.lines;
	const int nl = cl1 - cl0;
	if (nl <= 0) return; // warp-uniform

	constexpr int IB = IN16 ? 2 : 1, OB = (IN16 && !OUT8) ? 2 : 1;
	const long long in_pitch = pl.in_row_bytes, out_pitch = pl.out_row_bytes;
	const uint8_t* src = pl.in + (long long)t.f * p.in_frame_bytes + (long long)cl0 * in_pitch + (long long)k0 * IB;
	uint8_t* dst = pl.out + (long long)t.f * p.out_frame_bytes + (long long)cl0 * out_pitch + (long long)k0 * OB;

	const int b = active ? (k0 >> NSH) : 0; // idle lanes must not index past the register row
	const int i0 = k0 & (n - 1);
	GatherLane L;
	L.has_left = active && (i0 == 0) && (b > 0);
	L.has_right = active && (i0 + kSamplesPerLane == n) && (b + 1 < p.nb);
	const bool second_half = i0 != 0; // MERGE only
	const bool right_in_picture = k0 + kSamplesPerLane < pl.width; // samples right of the picture read as 0
	// who fetches a neighbour sample from memory: every lane in the host build, the warp's end lanes on the device
	const bool mem_left = L.has_left && (kHaloFromMemory || lane == 0);
	const bool mem_right = L.has_right && right_in_picture && (kHaloFromMemory || lane == 31);

	// vl/vr: neighbour samples fetched from memory (kept apart from raw: combining them would wait for the loads
	// just issued); MERGE keeps the lane's one neighbour in vl
	uint32_t raw[kGatherLB][4] = {}, vl[kGatherLB], vr[kGatherLB];
#pragma unroll
	for (int q = 0; q < kGatherLB; q++) {
		const uint8_t* row = src + (q < nl ? q : nl - 1) * in_pitch;
		if (active) {
			if (IN16) ld_global_16(row, raw[q]);
			else ld_global_8(row, raw[q]);
		}
		if (MERGE) {
			vl[q] = ld_sample_if<IB>(second_half ? row + kSamplesPerLane * IB : row - IB, mem_left || mem_right);
			vr[q] = 0;
		} else {
			vl[q] = ld_sample_if<IB>(row - IB, mem_left);
			vr[q] = ld_sample_if<IB>(row + kSamplesPerLane * IB, mem_right);
		}
	}

	const int bank = c ? 1 : 0;
	L.stride = p.pat_stride[bank];
	L.lut = luts + (smem_addr_t)(p.glut_index[c] * kLutBytes + lane * 4);
	L.pow16 = p.pow16;
	constexpr int kOutBias = (IN16 && OUT8) ? 2 : 0;
	L.lo2 = (uint32_t)(p.lo[c] + kOutBias) * 0x00010001u; L.hi2 = (uint32_t)(p.hi[c] + kOutBias) * 0x00010001u;

	const int srow = t.r - p.stream_row0;
	const uint16_t* w_cur = p.woffs + (((long long)t.f * p.stream_rows + srow) * p.spitch + 1 + b) * 4 + c;
	const smem_addr_t bank_addr = img + (smem_addr_t)p.gpat_off[bank];
	L.own = gather_window(bank_addr, w_cur[0], i0, L.s_own);
	L.lh = L.rh = L.own; L.s_l = L.s_r = 1;
	if (L.has_left) L.lh = gather_window(bank_addr, w_cur[-4], n - 1, L.s_l);
	if (L.has_right) (MERGE ? L.lh : L.rh) = gather_window(bank_addr, w_cur[4], 0, MERGE ? L.s_l : L.s_r);

	GatherUp U;
	U.own = U.lh = U.rh = L.own; U.s_own = U.s_l = U.s_r = 1;
	bool ovl = t.r > 0;
	if (ovl) {
		const uint16_t* w_up = w_cur - p.spitch * 4;
		U.own = gather_window(bank_addr, w_up[0], i0, U.s_own);
		if (L.has_left) U.lh = gather_window(bank_addr, w_up[-4], n - 1, U.s_l);
		if (L.has_right) (MERGE ? U.lh : U.rh) = gather_window(bank_addr, w_up[4], 0, MERGE ? U.s_l : U.s_r);
	}

	int rc = 0;
	const uint8_t* nxt = src + kGatherLB * in_pitch;
#pragma unroll 1
	for (int base = 0; base < nl; base += kGatherLB) {
#pragma unroll
		for (int q = 0; q < kGatherLB; q++) {
			const int line = base + q;
			uint32_t w[4];
			int w_cur = 0, w_up = 0, ru = 0;
			if (q == 0 && ovl) { w_cur = ysh ? 20 : 12; w_up = ysh ? 20 : 24; ru = (16 >> ysh) * L.stride; }
			if (q == 1 && ovl && !ysh) { w_cur = 24; w_up = 12; ru = 17 * L.stride; }
			uint32_t nl_s, nr_s = 0;
			if (MERGE) nl_s = neighbour_sample<IN16>(raw[q], lane, vl[q], second_half, right_in_picture);
			else neighbour_samples<IN16>(raw[q], lane, vl[q], vr[q], right_in_picture, nl_s, nr_s);
			gather_line<IN16, OUT8, MERGE>(L, U, rc, w_cur, w_up, ru, p.bs, raw[q], nl_s, nr_s, w);
			const bool more = line + kGatherLB < nl;
			if (IN16) ld_global_16_if(nxt, raw[q], more && active);
			else ld_global_8_if(nxt, raw[q], more && active);
			if (MERGE) {
				vl[q] = ld_sample_if<IB>(second_half ? nxt + kSamplesPerLane * IB : nxt - IB, (mem_left || mem_right) && more);
			} else {
				vl[q] = ld_sample_if<IB>(nxt - IB, mem_left && more);
				vr[q] = ld_sample_if<IB>(nxt + kSamplesPerLane * IB, mem_right && more);
			}
			if (line < nl && active) {
				if (OB == 2) st_global_16(dst, w);
				else st_global_8(dst, w);
			}
			rc += L.stride; nxt += in_pitch; dst += out_pitch;
		}
		ovl = false;
	}
}

// Dispatch on the component's block size (16 samples: luma and non-subsampled chroma; 8: chroma
// subsampled horizontally).
template <bool IN16, bool OUT8>
VFGS_HD void process_task_gather(const FgsParams& p, smem_addr_t luts, smem_addr_t img, uint32_t task, int lane)
{
	const TaskGeom t = decode_task(p, task);
	if (t.c && p.subx > 1) gather_task_body<IN16, OUT8, 3>(p, luts, img, t, lane);
	else gather_task_body<IN16, OUT8, 4>(p, luts, img, t, lane);
}

} // namespace vfgs
