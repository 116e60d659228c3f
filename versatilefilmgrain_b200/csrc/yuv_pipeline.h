// yuv_pipeline.h -- batched frame pipeline behind the reference's yuv.h interface (see include/yuv.h
// for the idea). Included at the end of vfgs_b200.cu: it shares the shim's state and error handling.
//
// Reference call sequence being served (src/vfgs_main.c:762-794):
//   yuv_alloc(frame) [yuv_alloc(oframe) if --outdepth 8]  yuv_skip
//   per frame: [vfgs_init_* when a cfg is due]  yuv_read  vfgs_add_grain -> vfgs_add_grain_line x H
//              [yuv_to_8bit]  yuv_write
//   yuv_free(frame) [yuv_free(oframe)]
//
// Three stages on three threads, linked by a ring of kPipeBatches batches of frame slots:
//   caller     yuv_read = fread straight into the batch being filled; yuv_write queues the frame
//   GPU stage  creates the CUDA context and page-locks the ring WHILE the caller already reads the first frames,
//              then takes every full batch through vfgs_b200_add_grain_frames_host (H2D, kernels, D2H)
//   writer     fwrite of the finished batches, in order, while the GPU stage works on the next batch
// Anything that touches the hardware state (vfgs_set_*, the end of the run) first drains the pipeline: queued frames
// were recorded under the state of their time.
#include <sys/mman.h>

#include <condition_variable>
#include <mutex>

namespace {

struct PipeSlot {
	bool grain = false; // vfgs_add_grain_line(y = 0) seen for this frame
	bool to8 = false;   // yuv_to_8bit seen
	FILE* out = nullptr;
};

constexpr int kPipeBatches = 4;
enum BatchState { kFree = 0, kFilling, kQueuedGpu, kQueuedWrite };

double pipe_now()
{
	timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}
const double g_lib_loaded = pipe_now();

struct Pipe {
	bool active = false;
	int w = 0, h = 0, depth = 0, fmt = 0;
	size_t ysz = 0, csz = 0;        // plane sizes in bytes at the input depth
	size_t in_frame = 0, out8_frame = 0;
	int cap = 0;                    // frames per batch
	uint8_t* ring = nullptr;        // kPipeBatches * cap * in_frame, page-locked by the GPU stage
	uint8_t* ring8 = nullptr;       // kPipeBatches * cap * out8_frame (only with --outdepth 8)
	size_t ring_bytes = 0, ring8_bytes = 0;
	bool ring_pinned = false, ring8_pinned = false;
	std::vector<PipeSlot> slot;     // kPipeBatches * cap
	int fill = 0;                   // batch being filled by the caller
	int cur = 0;                    // frames [0, cur) of that batch are queued, slot cur is being filled
	// hand-over between the stages
	std::mutex mu;
	std::condition_variable cv;
	BatchState state[kPipeBatches] = {kFree, kFree, kFree, kFree};
	int count[kPipeBatches] = {0, 0, 0, 0};
	int next_gpu = 0, next_write = 0; // batches are processed in the order they were filled
	bool quit = false;
	bool write_error = false;       // a short fwrite in the writer stage: reported by the next yuv_write
	std::thread gpu_thread, write_thread;
	unsigned long long frames_done = 0, flushes = 0;
	double t_read = 0, t_gpu = 0, t_write = 0; // seconds spent in fread, in the grain calls, in fwrite
	double t_ctx = 0, t_pin = 0, t_wait = 0;   // CUDA context creation, page-locking (both on the GPU stage, under the first reads), caller waiting
	~Pipe();
} g_pipe;

void pipe_at_exit();

bool pipe_owns(const void* p)
{
	const uint8_t* b = (const uint8_t*)p;
	return g_pipe.active && b >= g_pipe.ring && b < g_pipe.ring + g_pipe.ring_bytes;
}

void* pipe_map(size_t bytes)
{
	void* p = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
	return p == MAP_FAILED ? nullptr : p;
}

// GPU stage: grain synthesis of one batch, in frame order, one call per run of frames with the same requests
void pipe_gpu_batch(int b, int n)
{
	Pipe& P = g_pipe;
	PipeSlot* slot = &P.slot[(size_t)b * P.cap];
	uint8_t* ring = P.ring + (size_t)b * P.cap * P.in_frame;
	uint8_t* ring8 = P.ring8 ? P.ring8 + (size_t)b * P.cap * P.out8_frame : nullptr;
	if (P.ring8 && !P.ring8_pinned) { // the --outdepth 8 buffer is allocated after the input ring (src/vfgs_main.c:763-764)
		const double t0 = pipe_now();
		P.ring8_pinned = cudaHostRegister(P.ring8, P.ring8_bytes, cudaHostRegisterDefault) == cudaSuccess;
		P.t_pin += pipe_now() - t0;
	}
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
}

// writer stage: ordered writes, one fwrite per run of frames going to the same file
void pipe_write_batch(int b, int n)
{
	Pipe& P = g_pipe;
	PipeSlot* slot = &P.slot[(size_t)b * P.cap];
	const uint8_t* ring = P.ring + (size_t)b * P.cap * P.in_frame;
	const uint8_t* ring8 = P.ring8 ? P.ring8 + (size_t)b * P.cap * P.out8_frame : nullptr;
	for (int i = 0; i < n;) {
		int j = i + 1;
		while (j < n && slot[j].out == slot[i].out && slot[j].to8 == slot[i].to8) j++;
		if (slot[i].out) {
			const bool to8 = slot[i].to8;
			const size_t fb = to8 ? P.out8_frame : P.in_frame;
			const uint8_t* src = (to8 ? ring8 : ring) + (size_t)i * fb;
			if (fwrite(src, 1, fb * (size_t)(j - i), slot[i].out) != fb * (size_t)(j - i)) {
				fprintf(stderr, "vfgs_b200: short write (disk full?); yuv_write reports it from now on\n");
				std::lock_guard<std::mutex> lk(P.mu);
				P.write_error = true;
			}
		}
		i = j;
	}
}

void pipe_gpu_main()
{
	Pipe& P = g_pipe;
	// start-up, overlapped with the caller's first reads: CUDA context, then page-locking of the ring (the pages the
	// caller is writing meanwhile are simply pinned where they are)
	double t0 = pipe_now();
	if (ensure_ctx(-1)) fatal("frame pipeline");
	P.t_ctx = pipe_now() - t0;
	{ // after the CUDA runtime has registered its own exit handler: handlers run in reverse order, ours first
		static bool registered = false;
		if (!registered) { atexit(pipe_at_exit); registered = true; }
	}
	t0 = pipe_now();
	P.ring_pinned = cudaHostRegister(P.ring, P.ring_bytes, cudaHostRegisterDefault) == cudaSuccess; // pageable still works, slower
	P.t_pin += pipe_now() - t0;
	for (;;) {
		int b, n;
		{
			std::unique_lock<std::mutex> lk(P.mu);
			P.cv.wait(lk, [&] { return P.quit || P.state[P.next_gpu] == kQueuedGpu; });
			if (P.state[P.next_gpu] != kQueuedGpu) return;
			b = P.next_gpu; n = P.count[b];
		}
		t0 = pipe_now();
		pipe_gpu_batch(b, n);
		{
			std::lock_guard<std::mutex> lk(P.mu);
			P.t_gpu += pipe_now() - t0;
			P.state[b] = kQueuedWrite;
			P.next_gpu = (b + 1) % kPipeBatches;
		}
		P.cv.notify_all();
	}
}

void pipe_write_main()
{
	Pipe& P = g_pipe;
	for (;;) {
		int b, n;
		{
			std::unique_lock<std::mutex> lk(P.mu);
			P.cv.wait(lk, [&] { return P.quit || P.state[P.next_write] == kQueuedWrite; });
			if (P.state[P.next_write] != kQueuedWrite) return;
			b = P.next_write; n = P.count[b];
		}
		const double t0 = pipe_now();
		pipe_write_batch(b, n);
		{
			std::lock_guard<std::mutex> lk(P.mu);
			P.t_write += pipe_now() - t0;
			P.frames_done += (unsigned long long)n;
			P.flushes++;
			P.state[b] = kFree;
			P.next_write = (b + 1) % kPipeBatches;
		}
		P.cv.notify_all();
	}
}

// Hands the batch being filled to the GPU stage and continues in the next free one. wait: until everything queued
// has been written.
void pipe_flush(bool wait)
{
	Pipe& P = g_pipe;
	if (!P.active) return;
	const double t0 = pipe_now();
	std::unique_lock<std::mutex> lk(P.mu);
	if (P.cur > 0) {
		P.count[P.fill] = P.cur;
		P.state[P.fill] = kQueuedGpu;
		P.fill = (P.fill + 1) % kPipeBatches;
		P.cur = 0;
		P.cv.notify_all();
		P.cv.wait(lk, [&] { return P.state[P.fill] == kFree; }); // the oldest batch has been written
		P.state[P.fill] = kFilling;
	}
	if (wait)
		P.cv.wait(lk, [&] {
			for (int b = 0; b < kPipeBatches; b++)
				if (P.state[b] == kQueuedGpu || P.state[b] == kQueuedWrite) return false;
			return true;
		});
	P.t_wait += pipe_now() - t0;
}

void pipe_stop_threads()
{
	Pipe& P = g_pipe;
	{
		std::lock_guard<std::mutex> lk(P.mu);
		P.quit = true;
	}
	P.cv.notify_all();
	if (P.gpu_thread.joinable()) P.gpu_thread.join();
	if (P.write_thread.joinable()) P.write_thread.join();
	P.quit = false;
}

// A caller that exits without yuv_free (the reference's own main always calls it, src/vfgs_main.c:792) still gets its
// queued frames written: registered with atexit once the pipeline's context exists; the destructor only makes sure no
// thread outlives the object.
void pipe_at_exit()
{
	if (g_pipe.active) { pipe_flush(true); pipe_stop_threads(); }
}
Pipe::~Pipe()
{
	if (gpu_thread.joinable() || write_thread.joinable()) pipe_stop_threads();
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
	if (y == 0) g_pipe.slot[(size_t)g_pipe.fill * g_pipe.cap + g_pipe.cur].grain = true;
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

constexpr size_t kBatchBytes = 128u << 20; // input bytes buffered per batch before the GPU is fed

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
		// the input frame (src/vfgs_main.c:762): becomes the ring. Plain anonymous memory now; the GPU stage creates the
		// CUDA context and page-locks the ring while the caller is already reading frames into it.
		P.w = width; P.h = height; P.depth = depth; P.fmt = format;
		P.ysz = ysz; P.csz = csz;
		P.in_frame = ysz + 2 * csz;
		P.out8_frame = P.in_frame / sz;
		long long cap = (long long)(kBatchBytes / P.in_frame);
		P.cap = (int)(cap < 1 ? 1 : cap > 64 ? 64 : cap);
		P.ring_bytes = (size_t)kPipeBatches * P.cap * P.in_frame;
		P.ring = (uint8_t*)pipe_map(P.ring_bytes);
		if (!P.ring) return 1;
		P.slot.assign((size_t)kPipeBatches * P.cap, PipeSlot());
		for (int b = 0; b < kPipeBatches; b++) { P.state[b] = kFree; P.count[b] = 0; }
		P.fill = 0; P.cur = 0; P.next_gpu = P.next_write = 0;
		P.state[0] = kFilling;
		P.write_error = false;
		P.active = true;
		P.gpu_thread = std::thread(pipe_gpu_main);
		P.write_thread = std::thread(pipe_write_main);
		frame->Y = P.ring; frame->U = P.ring + ysz; frame->V = P.ring + ysz + csz;
		return 0;
	}
	// a second frame of the same geometry at depth 8 is the --outdepth 8 output (src/vfgs_main.c:763-764)
	if (width != P.w || height != P.h || format != P.fmt || depth != 8 || P.ring8) {
		snprintf(g_err, sizeof(g_err), "unexpected second yuv_alloc(%d, %d, %d, %d)", width, height, depth, format);
		fatal("yuv_alloc");
	}
	P.ring8_bytes = (size_t)kPipeBatches * P.cap * P.out8_frame;
	uint8_t* r8 = (uint8_t*)pipe_map(P.ring8_bytes);
	if (!r8) return 1;
	{
		std::lock_guard<std::mutex> lk(P.mu);
		P.ring8 = r8;
	}
	frame->Y = P.ring8; frame->U = P.ring8 + ysz; frame->V = P.ring8 + ysz + csz;
	return 0;
}

void yuv_free(yuv* frame)
{
	Pipe& P = g_pipe;
	pipe_flush(true); // src/vfgs_main.c:792: the end of the run drains the pipeline
	const bool is_ring = P.active && frame->Y && (pipe_owns(frame->Y) || frame->Y == P.ring);
	if (is_ring && getenv("VFGS_B200_PIPE_STATS"))
		fprintf(stderr, "vfgs_b200 pipeline: %llu frames, %llu batches; caller thread: fread %.3f s, waiting for a free batch %.3f s; GPU stage: "
		        "grain calls (H2D + kernels + D2H) %.3f s; writer stage: fwrite %.3f s; start-up on the GPU stage, under the first reads: CUDA context %.3f s, "
		        "page-locking %.3f s; %.3f s since library load\n",
		        P.frames_done, P.flushes, P.t_read, P.t_wait, P.t_gpu, P.t_write, P.t_ctx, P.t_pin, pipe_now() - g_lib_loaded);
	if (is_ring) {
		pipe_stop_threads();
		if (P.ring_pinned) cudaHostUnregister(P.ring);
		munmap(P.ring, P.ring_bytes);
		P.ring = nullptr; P.ring_pinned = false; P.active = false;
	} else if (P.ring8 && frame->Y == P.ring8) {
		if (P.ring8_pinned) cudaHostUnregister(P.ring8);
		munmap(P.ring8, P.ring8_bytes);
		P.ring8 = nullptr; P.ring8_pinned = false;
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
	uint8_t* p = P.ring + ((size_t)P.fill * P.cap + P.cur) * P.in_frame;
	frame->Y = p; frame->U = p + P.ysz; frame->V = p + P.ysz + P.csz; // next slot (the caller passes these on)
	P.slot[(size_t)P.fill * P.cap + P.cur] = PipeSlot();
	const double t0 = pipe_now();
	const bool bad = fread(p, 1, P.in_frame, file) != P.in_frame;
	P.t_read += pipe_now() - t0;
	return bad;
}

// Deferred: the frame is written by the writer stage, in order. Returns non-zero once an earlier deferred write has
// failed (the reference's yuv_write reports its own short write, src/yuv.c:207-214; here the report comes one or
// more frames late).
int yuv_write(yuv* frame, FILE* file)
{
	(void)frame; // the CLI's output struct holds stale or dummy pointers: the current slot is what is written
	Pipe& P = g_pipe;
	if (!P.active) return 1;
	P.slot[(size_t)P.fill * P.cap + P.cur].out = file;
	P.cur++;
	if (P.cur == P.cap) pipe_flush(false);
	std::lock_guard<std::mutex> lk(P.mu);
	return P.write_error ? 1 : 0;
}

void yuv_to_8bit(yuv* dst, const yuv* src)
{
	(void)dst;
	Pipe& P = g_pipe;
	if (!pipe_owns(src->Y) || !P.ring8 || P.depth != 10) {
		snprintf(g_err, sizeof(g_err), "yuv_to_8bit outside the frame pipeline");
		fatal("yuv_to_8bit");
	}
	P.slot[(size_t)P.fill * P.cap + P.cur].to8 = true; // (v + 2) >> 2 happens in the kernel's store
}

// pipeline statistics for tests: frames processed, batches written
void vfgs_b200_pipeline_stats(unsigned long long out[2])
{
	pipe_flush(true);
	std::lock_guard<std::mutex> lk(g_pipe.mu);
	out[0] = g_pipe.frames_done; out[1] = g_pipe.flushes;
}

} // extern "C"
