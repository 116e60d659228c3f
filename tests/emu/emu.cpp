// emu.cpp -- TEST TOOL, never part of the product library.
// Compiles the per-lane task code of the CUDA kernel (versatilefilmgrain_b200/csrc/fgs_task.h) with
// the HOST compiler and runs every (task, lane) pair serially, so the kernel's logic can be checked
// against the oracle in the GPU-less container before a GPU trip. It mirrors the shim's table-image
// and parameter construction (vfgs_b200.cu: build_blob / run_frames_device) from a dumped hw state.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "../../versatilefilmgrain_b200/csrc/fgs_task.h"

using namespace vfgs;

namespace {
struct StateDump { // == refh_state / oracle_dump
	int8_t pattern[2][9][64][64];
	uint8_t slut[3][256];
	uint8_t plut[3][256];
	uint32_t rnd, rnd_up, line_rnd, line_rnd_up;
	int scale_shift, bs, y_min, y_max, c_min, c_max, csubx, csuby;
};
} // namespace

extern "C" int emu_state_size(void) { return (int)sizeof(StateDump); }

// Packed planar frames, whole frames, like vfgs_b200_add_grain_frames_device. Returns 0.
extern "C" int emu_add_grain_frames(const void* state, const void* in, void* out, int nframes, int width,
                                    int height, int out_depth, int first_frame_index)
{
	const StateDump& h = *(const StateDump*)state;
	static const JumpTable jt;
	const int in_depth = 8 + h.bs;
	if (!out_depth) out_depth = in_depth;
	const int cw = width / h.csubx, ch = height / h.csuby;
	const int nb = (width + 15) / 16, R = (height + 15) / 16, wpr = ((nb + 31) >> 5) + 3;
	const size_t isz = in_depth > 8 ? 2 : 1, osz = out_depth > 8 ? 2 : 1;
	const size_t ysam = (size_t)width * height, csam = (size_t)cw * ch;

	// table image
	FgsParams p;
	memset(&p, 0, sizeof(p));
	int nslot[2] = {1, 1};
	for (int c = 0; c < 3; c++) {
		int first = h.plut[c][0] >> 4, uni = first;
		for (int i = 0; i < 256; i++) {
			int s = h.plut[c][i] >> 4;
			if (s != first) uni = -1;
			if (s + 1 > nslot[c ? 1 : 0]) nslot[c ? 1 : 0] = s + 1;
		}
		p.uniform_pi[c] = uni;
	}
	const int crows = 64 / h.csuby, ccols = 64 / h.csubx;
	p.lut_off = 0;
	p.pat_off[0] = 1536; p.pat_size[0] = 4096; p.pat_stride[0] = 64;
	p.pat_off[1] = 1536 + nslot[0] * 4096; p.pat_size[1] = crows * ccols; p.pat_stride[1] = ccols;
	p.blob_bytes = (p.pat_off[1] + nslot[1] * p.pat_size[1] + 31) & ~15;
	std::vector<uint8_t> blob((size_t)p.blob_bytes + 16, 0);
	uint8_t* tab = (uint8_t*)(((uintptr_t)blob.data() + 15) & ~(uintptr_t)15);
	uint16_t* lut = (uint16_t*)tab;
	for (int c = 0; c < 3; c++)
		for (int i = 0; i < 256; i++) lut[c * 256 + i] = (uint16_t)(h.slut[c][i] | ((h.plut[c][i] >> 4) << 8));
	for (int s = 0; s < nslot[0]; s++) memcpy(tab + p.pat_off[0] + s * 4096, h.pattern[0][s], 4096);
	for (int s = 0; s < nslot[1]; s++)
		for (int r = 0; r < crows; r++) memcpy(tab + p.pat_off[1] + s * p.pat_size[1] + r * ccols, h.pattern[1][s][r], (size_t)ccols);

	// LFSR streams (what lfsr_streams_kernel produces)
	std::vector<uint32_t> streams((size_t)nframes * R * wpr);
	for (int f = 0; f < nframes; f++)
		for (int r = 0; r < R; r++) {
			uint64_t t = ((uint64_t)(first_frame_index + f) * (uint64_t)(R - 1) + (uint64_t)r) * (uint64_t)nb;
			uint32_t s = jt.jump(h.line_rnd, t);
			for (int w = 0; w < wpr; w++) { streams[((size_t)f * R + r) * wpr + w] = s; s = jt.jump(s, 32); }
		}

	p.nframes = nframes; p.nb = nb; p.R = R; p.row_begin = 0; p.rows = R;
	p.y_begin = 0; p.y_end = height;
	p.subx = h.csubx; p.suby = h.csuby;
	p.in_bytes = (int)isz; p.out_bytes = (int)osz;
	p.bs = h.bs; p.ss = h.scale_shift;
	for (int c = 0; c < 3; c++) {
		p.lo[c] = (c ? h.c_min : h.y_min) << h.bs;
		p.hi[c] = (c ? h.c_max : h.y_max) << h.bs;
	}
	p.in_frame_bytes = (long long)((ysam + 2 * csam) * isz);
	p.out_frame_bytes = (long long)((ysam + 2 * csam) * osz);
	const size_t off[3] = {0, ysam, ysam + csam};
	for (int c = 0; c < 3; c++) {
		p.comp[c].in = (const uint8_t*)in + off[c] * isz;
		p.comp[c].out = (uint8_t*)out + off[c] * osz;
		p.comp[c].in_row_bytes = (long long)((c ? cw : width) * isz);
		p.comp[c].out_row_bytes = (long long)((c ? cw : width) * osz);
		p.comp[c].width = c ? cw : width;
		p.comp[c].lines = c ? ch : height;
		p.comp[c].vec = ((uintptr_t)p.comp[c].in % (8 * isz)) == 0 && (p.comp[c].in_row_bytes % (long long)(8 * isz)) == 0 &&
		                (p.in_frame_bytes % (long long)(8 * isz)) == 0 && ((uintptr_t)p.comp[c].out % (8 * osz)) == 0 &&
		                (p.comp[c].out_row_bytes % (long long)(8 * osz)) == 0 && (p.out_frame_bytes % (long long)(8 * osz)) == 0;
		p.nseg[c] = (p.comp[c].width + kSegSamples - 1) / kSegSamples;
	}
	p.tasks_per_stripe = p.nseg[0] + p.nseg[1] + p.nseg[2];
	p.total_tasks = (long long)nframes * p.rows * p.tasks_per_stripe;
	p.streams = streams.data(); p.wpr = wpr; p.stream_rows = R; p.stream_row0 = 0;

	for (long long task = 0; task < p.total_tasks; task++)
		for (int lane = 0; lane < 32; lane++) process_task(p, tab, task, lane);
	return 0;
}
