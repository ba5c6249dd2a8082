// p2p_probe.cu -- feasibility + cost of the peer-memory primitives the sharded path is built on (one process per GPU,
// buffers shared through CUDA IPC): cross-GPU flag barrier, cp.async.bulk (TMA engine) from PEER global memory into
// shared memory, random 8-byte peer reads, scattered 1-byte / 8-byte peer stores.  Not part of the product library.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); return -1; } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// every rank: write `epoch` into slot `me` of every peer's flag array, wait until all own slots reached it
__global__ void xbarrier_kernel(unsigned long long** flags, int me, int n, unsigned long long epoch, unsigned long long* err) {
    const int r = threadIdx.x;
    if (r < n) {
        __threadfence_system();
        *((volatile unsigned long long*)(flags[r] + me)) = epoch;
        const long long t0 = clock64();
        while (*((volatile unsigned long long*)(flags[me] + r)) < epoch) {
            if (clock64() - t0 > 4000000000ll) { *err = 1; break; }
        }
        __threadfence_system();
    }
}

// each CTA streams chunks of `chunk` bytes from src (local or peer) into shared memory with cp.async.bulk + mbarrier
__global__ void __launch_bounds__(128) bulk_pull_kernel(const uint8_t* src, uint64_t n_bytes, uint32_t chunk, int pieces, unsigned long long* sink) {
    extern __shared__ __align__(128) uint8_t sm[];
    __shared__ __align__(8) uint64_t bar[2];
    const uint32_t stage_bytes = chunk * pieces;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; i++) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar[i])), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const uint64_t n_stage = n_bytes / stage_bytes;
    unsigned long long acc = 0;
    uint32_t it = 0;
    auto issue = [&](uint64_t s, int st) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar[st])), "r"(stage_bytes) : "memory");
        for (int p = 0; p < pieces; p++) {
            // pieces come from `pieces` regions far apart (like one segment per sender)
            const uint64_t off = ((s * 2654435761ull) % n_stage) * stage_bytes / pieces / 16 * 16 + (uint64_t)p * (n_bytes / pieces);
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(sm + (size_t)st * stage_bytes + (size_t)p * chunk)),
                         "l"(src + (off % (n_bytes - chunk)) / 16 * 16), "r"(chunk), "r"(smem_u32(&bar[st]))
                         : "memory");
        }
    };
    uint64_t s = blockIdx.x;
    if (threadIdx.x == 0 && s < n_stage) issue(s, 0);
    for (; s < n_stage; s += gridDim.x, it++) {
        const int st = it & 1;
        if (threadIdx.x == 0 && s + gridDim.x < n_stage) issue(s + gridDim.x, st ^ 1);
        const uint32_t parity = (it >> 1) & 1;
        asm volatile(
            "{\n.reg .pred P1;\nW: mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra D;\nbra W;\nD:\n}" ::"r"(smem_u32(&bar[st])), "r"(parity) : "memory");
        const uint64_t* w = reinterpret_cast<const uint64_t*>(sm + (size_t)st * stage_bytes);
        for (uint32_t i = threadIdx.x; i < stage_bytes / 8; i += blockDim.x) acc += w[i];
        __syncthreads();
    }
    if (acc == 0x1234567ull) atomicAdd(sink, acc);
    if (threadIdx.x == 0) atomicAdd(sink + 1, acc);
}

__global__ void fill_kernel(uint64_t* p, uint64_t n, uint64_t v) { for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) p[i] = v + i; }
__global__ void rand_read8_kernel(const uint64_t* p, uint64_t n, uint64_t cnt, unsigned long long* sink) {
    unsigned long long a = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < cnt; i += (uint64_t)gridDim.x * blockDim.x) a += p[(i * 0x9E3779B97F4A7C15ull >> 20) % n];
    if (a == 77) atomicAdd(sink, a);
}
__global__ void rand_store1_kernel(uint8_t* p, uint64_t n, uint64_t cnt) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < cnt; i += (uint64_t)gridDim.x * blockDim.x) p[(i * 0x9E3779B97F4A7C15ull >> 20) % n] = (uint8_t)i;
}
// runs of 10 consecutive bytes at random places, one thread per run
__global__ void run_store1_kernel(uint8_t* p, uint64_t n, uint64_t runs) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < runs; i += (uint64_t)gridDim.x * blockDim.x) {
        uint8_t* q = p + (i * 0x9E3779B97F4A7C15ull >> 20) % (n - 16);
        for (int j = 0; j < 10; j++) q[j] = (uint8_t)(i + j);
    }
}
__global__ void rand_store8_kernel(uint64_t* p, uint64_t n, uint64_t cnt) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < cnt; i += (uint64_t)gridDim.x * blockDim.x) p[(i * 0x9E3779B97F4A7C15ull >> 20) % n] = i;
}
__global__ void stream_copy_kernel(const uint4* src, uint4* dst, uint64_t n16) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (uint64_t)gridDim.x * blockDim.x) dst[i] = src[i];
}

static float timed(cudaEvent_t a, cudaEvent_t b) { float ms = 0; cudaEventElapsedTime(&ms, a, b); return ms; }

extern "C" {
int p_alloc(int dev, uint64_t bytes, void** p, unsigned char* handle) {
    CK(cudaSetDevice(dev));
    CK(cudaMalloc(p, bytes));
    CK(cudaMemset(*p, 0, bytes));
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h, *p));
    memcpy(handle, &h, sizeof(h));
    return 0;
}
int p_open(int dev, const unsigned char* handle, void** p) {
    CK(cudaSetDevice(dev));
    cudaIpcMemHandle_t h; memcpy(&h, handle, sizeof(h));
    CK(cudaIpcOpenMemHandle(p, h, cudaIpcMemLazyEnablePeerAccess));
    return 0;
}
int p_fill(void* p, uint64_t bytes, uint64_t v) { fill_kernel<<<1024, 256>>>((uint64_t*)p, bytes / 8, v); CK(cudaDeviceSynchronize()); return 0; }
// flags_host: n device pointers (own + peers) to flag arrays of >= n u64
double p_barrier(void** flags_host, int me, int n, int reps, unsigned long long epoch0) {
    unsigned long long** d_flags; unsigned long long* d_err;
    if (cudaMalloc(&d_flags, n * sizeof(void*)) != cudaSuccess) return -1;
    cudaMalloc(&d_err, 8); cudaMemset(d_err, 0, 8);
    cudaMemcpy(d_flags, flags_host, n * sizeof(void*), cudaMemcpyHostToDevice);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    xbarrier_kernel<<<1, 32>>>(d_flags, me, n, epoch0, d_err);
    cudaEventRecord(a);
    for (int i = 1; i <= reps; i++) xbarrier_kernel<<<1, 32>>>(d_flags, me, n, epoch0 + i, d_err);
    cudaEventRecord(b);
    if (cudaDeviceSynchronize() != cudaSuccess) return -2;
    unsigned long long err = 0; cudaMemcpy(&err, d_err, 8, cudaMemcpyDeviceToHost);
    if (err) return -3;
    return timed(a, b) / reps * 1000.0;  // us per barrier
}
// GB/s of cp.async.bulk pulls from `src`
double p_bulk(const void* src, uint64_t n_bytes, uint32_t chunk, int pieces, int ctas) {
    unsigned long long* sink; cudaMalloc(&sink, 16); cudaMemset(sink, 0, 16);
    const size_t smem = (size_t)2 * chunk * pieces;
    if (cudaFuncSetAttribute(bulk_pull_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    bulk_pull_kernel<<<ctas, 128, smem>>>((const uint8_t*)src, n_bytes, chunk, pieces, sink);
    cudaEventRecord(a);
    bulk_pull_kernel<<<ctas, 128, smem>>>((const uint8_t*)src, n_bytes, chunk, pieces, sink);
    cudaEventRecord(b);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("bulk: %s\n", cudaGetErrorString(e)); return -2; }
    const uint64_t moved = n_bytes / ((uint64_t)chunk * pieces) * ((uint64_t)chunk * pieces);
    return moved / (timed(a, b) * 1e-3) / 1e9;
}
double p_bulk_sum(const void* src, uint64_t n_bytes, uint32_t chunk, int pieces, int ctas) {  // checksum of one pass (correctness local vs peer)
    unsigned long long* sink; cudaMalloc(&sink, 16); cudaMemset(sink, 0, 16);
    const size_t smem = (size_t)2 * chunk * pieces;
    cudaFuncSetAttribute(bulk_pull_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    bulk_pull_kernel<<<ctas, 128, smem>>>((const uint8_t*)src, n_bytes, chunk, pieces, sink);
    if (cudaDeviceSynchronize() != cudaSuccess) return -1;
    unsigned long long h[2]; cudaMemcpy(h, sink, 16, cudaMemcpyDeviceToHost);
    return (double)(h[1] & 0xffffffffffffull);
}
// mode 0: random 8 B reads, 1: random 1 B stores, 2: runs of 10 x 1 B stores, 3: random 8 B stores, 4: streaming copy src->dst (dst local)
double p_rate(int mode, void* p, uint64_t bytes, uint64_t cnt, void* local_dst) {
    unsigned long long* sink; cudaMalloc(&sink, 8); cudaMemset(sink, 0, 8);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int rep = 0; rep < 2; rep++) {
        if (rep == 1) cudaEventRecord(a);
        if (mode == 0) rand_read8_kernel<<<148 * 16, 256>>>((const uint64_t*)p, bytes / 8, cnt, sink);
        else if (mode == 1) rand_store1_kernel<<<148 * 16, 256>>>((uint8_t*)p, bytes, cnt);
        else if (mode == 2) run_store1_kernel<<<148 * 16, 256>>>((uint8_t*)p, bytes, cnt);
        else if (mode == 3) rand_store8_kernel<<<148 * 16, 256>>>((uint64_t*)p, bytes / 8, cnt);
        else stream_copy_kernel<<<148 * 16, 256>>>((const uint4*)p, (uint4*)local_dst, bytes / 16);
    }
    cudaEventRecord(b);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("rate mode %d: %s\n", mode, cudaGetErrorString(e)); return -1; }
    return timed(a, b);  // ms
}
}
