// yuv_pipeline.h -- batched frame pipeline behind the reference's yuv.h interface (see include/yuv.h
// for the idea). Included at the end of vfgs_b200.cu: it shares the shim's state and error handling.
//
// Reference call sequence being served (src/vfgs_main.c:762-794):
//   yuv_alloc(frame) [yuv_alloc(oframe) if --outdepth 8]  yuv_skip
//   per frame: [vfgs_init_* when a cfg is due]  yuv_read  vfgs_add_grain -> vfgs_add_grain_line x H
//              [yuv_to_8bit]  yuv_write
//   yuv_free(frame) [yuv_free(oframe)]

namespace {

struct PipeSlot {
	bool grain = false; // vfgs_add_grain_line(y = 0) seen for this frame
	bool to8 = false;   // yuv_to_8bit seen
	FILE* out = nullptr;
};

struct Pipe {
	bool active = false;
	int w = 0, h = 0, depth = 0, fmt = 0;
	size_t ysz = 0, csz = 0;        // plane sizes in bytes at the input depth
	size_t in_frame = 0, out8_frame = 0;
	int cap = 0;                    // ring capacity in frames
	uint8_t* ring = nullptr;        // page-locked, cap * in_frame
	uint8_t* ring8 = nullptr;       // page-locked, cap * out8_frame (only with --outdepth 8)
	std::vector<PipeSlot> slot;
	int cur = 0;                    // frames [0, cur) are queued, slot cur is being filled
	unsigned long long frames_done = 0, flushes = 0;
} g_pipe;

constexpr size_t kRingBytes = 256u << 20; // input bytes buffered before the GPU is fed

bool pipe_owns(const void* p)
{
	const uint8_t* b = (const uint8_t*)p;
	return g_pipe.active && b >= g_pipe.ring && b < g_pipe.ring + (size_t)g_pipe.cap * g_pipe.in_frame;
}

void pipe_flush()
{
	Pipe& P = g_pipe;
	if (!P.active || P.cur == 0) return;
	const int n = P.cur;
	P.cur = 0; // setters called from inside must not recurse
	for (int i = 0; i < n;) {
		int j = i + 1;
		while (j < n && P.slot[j].grain == P.slot[i].grain && P.slot[j].to8 == P.slot[i].to8) j++;
		if (!P.slot[i].grain) {
			snprintf(g_err, sizeof(g_err), "yuv_write of a frame that never went through vfgs_add_grain");
			fatal("frame pipeline");
		}
		const bool to8 = P.slot[i].to8;
		uint8_t* in = P.ring + (size_t)i * P.in_frame;
		uint8_t* out = to8 ? P.ring8 + (size_t)i * P.out8_frame : in;
		if (vfgs_b200_add_grain_frames_host(in, out, j - i, P.w, P.h, to8 ? 8 : 0) != VFGS_B200_OK) fatal("frame pipeline");
		i = j;
	}
	// ordered writes, one fwrite per run of frames going to the same file
	for (int i = 0; i < n;) {
		int j = i + 1;
		while (j < n && P.slot[j].out == P.slot[i].out && P.slot[j].to8 == P.slot[i].to8) j++;
		if (P.slot[i].out) {
			const bool to8 = P.slot[i].to8;
			const size_t fb = to8 ? P.out8_frame : P.in_frame;
			const uint8_t* src = (to8 ? P.ring8 : P.ring) + (size_t)i * fb;
			if (fwrite(src, 1, fb * (size_t)(j - i), P.slot[i].out) != fb * (size_t)(j - i))
				fprintf(stderr, "vfgs_b200: short write\n");
		}
		i = j;
	}
	P.frames_done += (unsigned long long)n;
	P.flushes++;
}

// Hook of every vfgs_set_*: queued frames were recorded under the current state.
void pipe_before_state_change()
{
	if (g_pipe.active && g_pipe.cur > 0) pipe_flush();
}

// Hook of vfgs_add_grain_line: true when the line belongs to a pipeline slot (nothing to do per line).
bool pipe_line(const void* Y, int y)
{
	if (!pipe_owns(Y)) return false;
	if (y == 0) g_pipe.slot[g_pipe.cur].grain = true;
	return true;
}

void fill_yuv(yuv* f, int width, int height, int depth, int format)
{
	const int subx = format > YUV_422 ? 1 : 2, suby = format > YUV_420 ? 1 : 2; // yuv.c:59-60
	f->depth = (unsigned)depth;
	f->width = (unsigned short)width; f->height = (unsigned short)height;
	f->stride = (unsigned short)width;               // packed: the file layout is the buffer layout
	f->cwidth = (unsigned short)(width / subx); f->cheight = (unsigned short)(height / suby);
	f->cstride = f->cwidth;
}

} // namespace

extern "C" {

int yuv_alloc(int width, int height, int depth, int format, yuv* frame)
{
	Pipe& P = g_pipe;
	fill_yuv(frame, width, height, depth, format);
	frame->Y = frame->U = frame->V = nullptr;
	const size_t sz = depth > 8 ? 2 : 1;
	const size_t ysz = (size_t)width * height * sz, csz = (size_t)frame->cwidth * frame->cheight * sz;
	if (!P.active) {
		// the input frame (src/vfgs_main.c:762): becomes the ring
		if (ensure_ctx(-1)) fatal("yuv_alloc");
		P.w = width; P.h = height; P.depth = depth; P.fmt = format;
		P.ysz = ysz; P.csz = csz;
		P.in_frame = ysz + 2 * csz;
		P.out8_frame = P.in_frame / sz;
		long long cap = (long long)(kRingBytes / P.in_frame);
		P.cap = (int)(cap < 2 ? 2 : cap > 64 ? 64 : cap);
		P.ring = (uint8_t*)vfgs_b200_host_alloc((size_t)P.cap * P.in_frame);
		if (!P.ring) return 1;
		P.slot.assign((size_t)P.cap, PipeSlot());
		P.cur = 0;
		P.active = true;
		frame->Y = P.ring; frame->U = P.ring + ysz; frame->V = P.ring + ysz + csz;
		return 0;
	}
	// a second frame of the same geometry at depth 8 is the --outdepth 8 output (src/vfgs_main.c:763-764)
	if (width != P.w || height != P.h || format != P.fmt || depth != 8 || P.ring8) {
		snprintf(g_err, sizeof(g_err), "unexpected second yuv_alloc(%d, %d, %d, %d)", width, height, depth, format);
		fatal("yuv_alloc");
	}
	P.ring8 = (uint8_t*)vfgs_b200_host_alloc((size_t)P.cap * P.out8_frame);
	if (!P.ring8) return 1;
	frame->Y = P.ring8; frame->U = P.ring8 + ysz; frame->V = P.ring8 + ysz + csz;
	return 0;
}

void yuv_free(yuv* frame)
{
	Pipe& P = g_pipe;
	pipe_flush(); // src/vfgs_main.c:792: the end of the run drains the pipeline
	if (P.active && frame->Y && (pipe_owns(frame->Y) || frame->Y == P.ring)) {
		vfgs_b200_host_free(P.ring);
		P.ring = nullptr; P.active = false;
	} else if (P.ring8 && frame->Y == P.ring8) {
		vfgs_b200_host_free(P.ring8);
		P.ring8 = nullptr;
	}
	frame->Y = frame->U = frame->V = nullptr;
}

void yuv_pad(yuv* frame) { (void)frame; /* packed slots have no stride padding to fill */ }

int yuv_skip(yuv* frame, int n, FILE* file)
{
	const long long sz = frame->depth == 8 ? 1 : 2;
	const long long size = ((long long)frame->width * frame->height + 2ll * frame->cwidth * frame->cheight) * sz;
	return fseeko(file, (off_t)(size * n), SEEK_CUR);
}

int yuv_read(yuv* frame, FILE* file)
{
	Pipe& P = g_pipe;
	if (!P.active) return 1;
	if (P.cur == P.cap) pipe_flush();
	uint8_t* p = P.ring + (size_t)P.cur * P.in_frame;
	frame->Y = p; frame->U = p + P.ysz; frame->V = p + P.ysz + P.csz; // next slot (the caller passes these on)
	P.slot[P.cur] = PipeSlot();
	return fread(p, 1, P.in_frame, file) != P.in_frame;
}

int yuv_write(yuv* frame, FILE* file)
{
	(void)frame; // the CLI's output struct holds stale or dummy pointers: the current slot is what is written
	Pipe& P = g_pipe;
	if (!P.active) return 1;
	P.slot[P.cur].out = file;
	P.cur++;
	if (P.cur == P.cap) pipe_flush();
	return 0;
}

void yuv_to_8bit(yuv* dst, const yuv* src)
{
	(void)dst;
	Pipe& P = g_pipe;
	if (!pipe_owns(src->Y) || !P.ring8 || P.depth != 10) {
		snprintf(g_err, sizeof(g_err), "yuv_to_8bit outside the frame pipeline");
		fatal("yuv_to_8bit");
	}
	P.slot[P.cur].to8 = true; // (v + 2) >> 2 happens in the kernel's store
}

// pipeline statistics for tests: frames processed, flushes
void vfgs_b200_pipeline_stats(unsigned long long out[2])
{
	out[0] = g_pipe.frames_done; out[1] = g_pipe.flushes;
}

} // extern "C"
