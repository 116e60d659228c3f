// tma_ragged_copy.cu -- microbenchmark behind the TMA variant of the ragged-width grain kernel:
// can 1-D tensor-map copies with 512-byte boxes at arbitrary ELEMENT offsets (rows that are not 16-byte aligned)
// stream a picture through shared memory at HBM speed?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o build/tma_ragged_copy tools/tma_ragged_copy.cu
//   build/tma_ragged_copy [width] [lines] [slots]
// Every warp walks tasks of 16 lines x 256 samples (uint16): per line one TMA load into a ring slot, one 16-byte
// shared load per lane, +1, one 16-byte shared store, one TMA store. The last task of a row is shorter (second map).
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(n)); }
__device__ __forceinline__ void mbar_expect(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
	asm volatile("{\n\t.reg .pred p;\n\tW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@!p bra W;\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_1d(uint32_t dst, const CUtensorMap* map, int c0, uint32_t bar)
{
	// the buffer is described as OVERLAPPING rows of 264 samples, 16 bytes apart (make_map): sample e = row e / 8, column e % 8
	asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
	             ::"r"(dst), "l"((uint64_t)map), "r"(c0 & 7), "r"(c0 >> 3), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma_store_1d(const CUtensorMap* map, int c0, uint32_t src)
{
	asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"((uint64_t)map), "r"(c0 & 7), "r"(c0 >> 3), "r"(src) : "memory");
}

constexpr int kWarps = 24, kLines = 16;

template <int S>
__global__ void __launch_bounds__(kWarps * 32) copy_kernel(const __grid_constant__ CUtensorMap in_map, const __grid_constant__ CUtensorMap out_map,
                                                            const __grid_constant__ CUtensorMap out_tail, int mode, int width, int lines, int segs, int tail_units, long long tasks)
{
	extern __shared__ __align__(128) uint8_t smem[];
	__shared__ __align__(8) uint64_t bars[kWarps][S];
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	uint8_t* ring_in = smem + (size_t)warp * S * 1024;
	uint8_t* ring_out = ring_in + S * 512;
	if (lane == 0) for (int s = 0; s < S; s++) mbar_init(smem_u32(&bars[warp][s]), 1);
	asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	__syncwarp();
	uint32_t phase = 0; // bit s: parity of slot s
	const int lmask = (mode & 1) ? ~7 : ~0, smask = (mode & 2) ? ~7 : ~0; // experiments: force aligned loads / stores (wrong data, timing only)
	const int stripes = lines / kLines;
	for (long long t = (long long)blockIdx.x * kWarps + warp; t < tasks; t += (long long)gridDim.x * kWarps) {
		const int seg = (int)(t % segs);
		const long long rest = t / segs;
		const int stripe = (int)(rest % stripes);
		const long long plane = rest / stripes;
		const long long e0 = (plane * lines + (long long)stripe * kLines) * width + seg * 256;
		const bool tail = seg == segs - 1 && tail_units < 32;
		if (lane == 0)
			for (int s = 0; s < S; s++) {
				mbar_expect(smem_u32(&bars[warp][s]), 512);
				tma_load_1d(smem_u32(ring_in + s * 512), &in_map, (int)(e0 + (long long)s * width) & lmask, smem_u32(&bars[warp][s]));
			}
#pragma unroll 1
		for (int l = 0; l < kLines; l++) {
			const int s = l % S;
			mbar_wait(smem_u32(&bars[warp][s]), (phase >> s) & 1);
			phase ^= 1u << s;
			uint4 v = *(const uint4*)(ring_in + s * 512 + lane * 16);
			v.x += 0x00010001u; v.y += 0x00010001u; v.z += 0x00010001u; v.w += 0x00010001u;
			if (lane == 0) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(S - 1) : "memory"); // the store that last read this out slot
			__syncwarp();
			*(uint4*)(ring_out + s * 512 + lane * 16) = v;
			asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
			__syncwarp();
			if (lane == 0) {
				tma_store_1d(tail ? &out_tail : &out_map, (int)(e0 + (long long)l * width) & smask, smem_u32(ring_out + s * 512));
				asm volatile("cp.async.bulk.commit_group;" ::: "memory");
				if (l + S < kLines) {
					mbar_expect(smem_u32(&bars[warp][s]), 512);
					tma_load_1d(smem_u32(ring_in + s * 512), &in_map, (int)(e0 + (long long)(l + S) * width) & lmask, smem_u32(&bars[warp][s]));
				}
			}
		}
	}
	if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                             CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static CUtensorMap make_map(EncodeFn enc, void* base, uint64_t elems, uint32_t box)
{
	CUtensorMap m;
	cuuint64_t dims[2] = {264, (elems + 7) / 8};
	cuuint64_t strides[1] = {16};
	cuuint32_t boxd[2] = {box, 1}, es[2] = {1, 1};
	CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, base, dims, strides, boxd, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
	                 CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
	if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d (box %u)\n", (int)r, box); exit(1); }
	return m;
}

static int g_mode = 0;
template <int S>
static void run(EncodeFn enc, int width, int lines, long long planes, uint16_t* din, uint16_t* dout)
{
	const uint64_t elems = (uint64_t)planes * lines * width;
	const int units = (width + 7) / 8, segs = (units + 31) / 32;
	int tail_units = (width / 8) - (segs - 1) * 32; // whole units of the last segment (a trailing partial unit is left out here)
	if (tail_units > 32) tail_units = 32;
	CUtensorMap mi = make_map(enc, din, elems, 256), mo = make_map(enc, dout, elems, 256), mt = make_map(enc, dout, elems, tail_units > 0 ? tail_units * 8 : 8);
	const long long tasks = planes * (lines / kLines) * segs;
	const size_t smem = (size_t)kWarps * S * 1024;
	CK(cudaFuncSetAttribute(copy_kernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
	cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
	for (int it = 0; it < 3; it++) copy_kernel<S><<<148, kWarps * 32, smem>>>(mi, mo, mt, g_mode, width, lines, segs, tail_units, tasks);
	CK(cudaDeviceSynchronize());
	const int reps = 10;
	CK(cudaEventRecord(a));
	for (int it = 0; it < reps; it++) copy_kernel<S><<<148, kWarps * 32, smem>>>(mi, mo, mt, g_mode, width, lines, segs, tail_units, tasks);
	CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
	float ms; CK(cudaEventElapsedTime(&ms, a, b));
	const double bytes = 2.0 * (double)planes * lines * (double)(width / 8 * 8) * 2.0;
	printf("width %d slots %d: %.1f us/launch, %.0f GB/s (read + write of whole units)\n", width, S, ms * 1e3 / reps, bytes * reps / (ms * 1e-3) / 1e9);
	// check a few planes
	std::vector<uint16_t> hi((size_t)lines * width), ho((size_t)lines * width);
	CK(cudaMemcpy(hi.data(), din, hi.size() * 2, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(ho.data(), dout, ho.size() * 2, cudaMemcpyDeviceToHost));
	long long bad = 0;
	for (int y = 0; y < lines; y++) for (int x = 0; x < width / 8 * 8; x++) if ((uint16_t)(hi[(size_t)y * width + x] + 1) != ho[(size_t)y * width + x]) bad++;
	printf("  first plane: %lld wrong samples\n", bad);
}

int main(int argc, char** argv)
{
	const int width = argc > 1 ? atoi(argv[1]) : 1366, lines = argc > 2 ? atoi(argv[2]) : 768, slots = argc > 3 ? atoi(argv[3]) : 4;
	g_mode = argc > 4 ? atoi(argv[4]) : 0;
	EncodeFn enc = nullptr; cudaDriverEntryPointQueryResult q;
	CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &q));
	if (!enc) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
	const long long planes = (long long)(1.0e9 / ((double)width * lines * 2)); // ~1 GB each way
	uint16_t *din, *dout;
	const size_t n = (size_t)planes * lines * width;
	CK(cudaMalloc(&din, n * 2 + 1024)); CK(cudaMalloc(&dout, n * 2 + 1024));
	std::vector<uint16_t> h((size_t)lines * width);
	for (size_t i = 0; i < h.size(); i++) h[i] = (uint16_t)((i * 2654435761u >> 20) & 0x3ff);
	for (long long p = 0; p < planes; p++) CK(cudaMemcpy(din + (size_t)p * h.size(), h.data(), h.size() * 2, cudaMemcpyHostToDevice));
	CK(cudaMemset(dout, 0, n * 2));
	if (slots == 2) run<2>(enc, width, lines, planes, din, dout);
	else if (slots == 3) run<3>(enc, width, lines, planes, din, dout);
	else if (slots == 6) run<6>(enc, width, lines, planes, din, dout);
	else if (slots == 8) run<8>(enc, width, lines, planes, din, dout);
	else run<4>(enc, width, lines, planes, din, dout);
	return 0;
}
