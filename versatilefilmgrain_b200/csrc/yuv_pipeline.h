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

// The ring is two halves: the caller's thread fills one (yuv_read = fread into page-locked memory) while a
// worker thread takes the other through the GPU (H2D, kernels, D2H) and writes it out, so file input overlaps
// grain synthesis and file output. Anything that touches the hardware state (vfgs_set_*, the end of the run)
// first waits for the worker: queued frames were recorded under the state of their time.
struct Pipe {
	bool active = false;
	int w = 0, h = 0, depth = 0, fmt = 0;
	size_t ysz = 0, csz = 0;        // plane sizes in bytes at the input depth
	size_t in_frame = 0, out8_frame = 0;
	int cap = 0;                    // frames per half
	uint8_t* ring = nullptr;        // page-locked, 2 * cap * in_frame
	uint8_t* ring8 = nullptr;       // page-locked, 2 * cap * out8_frame (only with --outdepth 8)
	std::vector<PipeSlot> slot;     // 2 * cap
	int half = 0;                   // half being filled
	int cur = 0;                    // frames [0, cur) of that half are queued, slot cur is being filled
	std::thread worker;             // processes the other half
	unsigned long long frames_done = 0, flushes = 0;
	double t_read = 0, t_gpu = 0, t_write = 0; // seconds spent in fread, in the grain calls, in fwrite
	double t_ctx = 0, t_alloc = 0, t_wait = 0; // CUDA context creation, page-locked allocation, caller waiting for the worker
	~Pipe() { if (worker.joinable()) worker.join(); } // a caller that exits without yuv_free must not trip std::terminate
} g_pipe;

double pipe_now()
{
	timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}
const double g_lib_loaded = pipe_now();

constexpr size_t kRingBytes = 128u << 20; // input bytes buffered per half before the GPU is fed

bool pipe_owns(const void* p)
{
	const uint8_t* b = (const uint8_t*)p;
	return g_pipe.active && b >= g_pipe.ring && b < g_pipe.ring + (size_t)2 * g_pipe.cap * g_pipe.in_frame;
}

// Grain synthesis and ordered output of frames [0, n) of one half (runs on the worker thread).
void pipe_process(int half, int n)
{
	Pipe& P = g_pipe;
	PipeSlot* slot = &P.slot[(size_t)half * P.cap];
	uint8_t* ring = P.ring + (size_t)half * P.cap * P.in_frame;
	uint8_t* ring8 = P.ring8 ? P.ring8 + (size_t)half * P.cap * P.out8_frame : nullptr;
	const double t0 = pipe_now();
	for (int i = 0; i < n;) {
		int j = i + 1;
		while (j < n && slot[j].grain == slot[i].grain && slot[j].to8 == slot[i].to8) j++;
		if (!slot[i].grain) {
			snprintf(g_err, sizeof(g_err), "yuv_write of a frame that never went through vfgs_add_grain");
			fatal("frame pipeline");
		}
		const bool to8 = slot[i].to8;
		uint8_t* in = ring + (size_t)i * P.in_frame;
		uint8_t* out = to8 ? ring8 + (size_t)i * P.out8_frame : in;
		if (vfgs_b200_add_grain_frames_host(in, out, j - i, P.w, P.h, to8 ? 8 : 0) != VFGS_B200_OK) fatal("frame pipeline");
		i = j;
	}
	const double t1 = pipe_now();
	// ordered writes, one fwrite per run of frames going to the same file
	for (int i = 0; i < n;) {
		int j = i + 1;
		while (j < n && slot[j].out == slot[i].out && slot[j].to8 == slot[i].to8) j++;
		if (slot[i].out) {
			const bool to8 = slot[i].to8;
			const size_t fb = to8 ? P.out8_frame : P.in_frame;
			const uint8_t* src = (to8 ? ring8 : ring) + (size_t)i * fb;
			if (fwrite(src, 1, fb * (size_t)(j - i), slot[i].out) != fb * (size_t)(j - i))
				fprintf(stderr, "vfgs_b200: short write\n");
		}
		i = j;
	}
	P.t_gpu += t1 - t0; P.t_write += pipe_now() - t1;
	P.frames_done += (unsigned long long)n;
	P.flushes++;
}

void pipe_wait()
{
	const double t0 = pipe_now();
	if (g_pipe.worker.joinable()) g_pipe.worker.join();
	g_pipe.t_wait += pipe_now() - t0;
}

// Hands the half being filled to the worker and continues in the other one. wait: also wait for that work.
void pipe_flush(bool wait)
{
	Pipe& P = g_pipe;
	if (!P.active) return;
	pipe_wait(); // the other half is free again once its job is done
	if (P.cur > 0) {
		const int half = P.half, n = P.cur;
		P.half ^= 1; P.cur = 0;
		P.worker = std::thread(pipe_process, half, n);
	}
	if (wait) pipe_wait();
}

// Hook of every vfgs_set_*: queued frames were recorded under the current state.
void pipe_before_state_change()
{
	if (g_pipe.active) pipe_flush(true);
}

// Hook of vfgs_add_grain_line: true when the line belongs to a pipeline slot (nothing to do per line).
bool pipe_line(const void* Y, int y)
{
	if (!pipe_owns(Y)) return false;
	if (y == 0) g_pipe.slot[(size_t)g_pipe.half * g_pipe.cap + g_pipe.cur].grain = true;
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
		const double t0 = pipe_now();
		if (ensure_ctx(-1)) fatal("yuv_alloc");
		P.t_ctx = pipe_now() - t0;
		P.w = width; P.h = height; P.depth = depth; P.fmt = format;
		P.ysz = ysz; P.csz = csz;
		P.in_frame = ysz + 2 * csz;
		P.out8_frame = P.in_frame / sz;
		long long cap = (long long)(kRingBytes / P.in_frame);
		P.cap = (int)(cap < 1 ? 1 : cap > 64 ? 64 : cap);
		const double t1 = pipe_now();
		P.ring = (uint8_t*)vfgs_b200_host_alloc((size_t)2 * P.cap * P.in_frame);
		if (!P.ring) return 1;
		P.t_alloc += pipe_now() - t1;
		P.slot.assign((size_t)2 * P.cap, PipeSlot());
		P.half = 0; P.cur = 0;
		P.active = true;
		frame->Y = P.ring; frame->U = P.ring + ysz; frame->V = P.ring + ysz + csz;
		return 0;
	}
	// a second frame of the same geometry at depth 8 is the --outdepth 8 output (src/vfgs_main.c:763-764)
	if (width != P.w || height != P.h || format != P.fmt || depth != 8 || P.ring8) {
		snprintf(g_err, sizeof(g_err), "unexpected second yuv_alloc(%d, %d, %d, %d)", width, height, depth, format);
		fatal("yuv_alloc");
	}
	P.ring8 = (uint8_t*)vfgs_b200_host_alloc((size_t)2 * P.cap * P.out8_frame);
	if (!P.ring8) return 1;
	frame->Y = P.ring8; frame->U = P.ring8 + ysz; frame->V = P.ring8 + ysz + csz;
	return 0;
}

void yuv_free(yuv* frame)
{
	Pipe& P = g_pipe;
	pipe_flush(true); // src/vfgs_main.c:792: the end of the run drains the pipeline
	if (P.active && getenv("VFGS_B200_PIPE_STATS") && frame->Y && (pipe_owns(frame->Y) || frame->Y == P.ring))
		fprintf(stderr, "vfgs_b200 pipeline: %llu frames, %llu flushes; caller thread: fread %.3f s, waiting for the worker %.3f s; worker thread: "
		        "grain calls (H2D + kernels + D2H) %.3f s, fwrite %.3f s; start-up: CUDA context %.3f s, page-locked ring %.3f s; %.3f s since library load\n",
		        P.frames_done, P.flushes, P.t_read, P.t_wait, P.t_gpu, P.t_write, P.t_ctx, P.t_alloc, pipe_now() - g_lib_loaded);
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
	if (P.cur == P.cap) pipe_flush(false);
	uint8_t* p = P.ring + ((size_t)P.half * P.cap + P.cur) * P.in_frame;
	frame->Y = p; frame->U = p + P.ysz; frame->V = p + P.ysz + P.csz; // next slot (the caller passes these on)
	P.slot[(size_t)P.half * P.cap + P.cur] = PipeSlot();
	const double t0 = pipe_now();
	const bool bad = fread(p, 1, P.in_frame, file) != P.in_frame;
	P.t_read += pipe_now() - t0;
	return bad;
}

int yuv_write(yuv* frame, FILE* file)
{
	(void)frame; // the CLI's output struct holds stale or dummy pointers: the current slot is what is written
	Pipe& P = g_pipe;
	if (!P.active) return 1;
	P.slot[(size_t)P.half * P.cap + P.cur].out = file;
	P.cur++;
	if (P.cur == P.cap) pipe_flush(false);
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
	P.slot[(size_t)P.half * P.cap + P.cur].to8 = true; // (v + 2) >> 2 happens in the kernel's store
}

// pipeline statistics for tests: frames processed, flushes
void vfgs_b200_pipeline_stats(unsigned long long out[2])
{
	pipe_wait();
	out[0] = g_pipe.frames_done; out[1] = g_pipe.flushes;
}

} // extern "C"
