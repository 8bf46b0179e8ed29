// Shared device/host helpers for the ProxyTTA B200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

typedef __nv_bfloat16 bf16;
typedef __nv_bfloat162 bf162;

namespace ptta {

// ---- error plumbing: every C-ABI entry returns 0 on success, non-zero otherwise -----------------
void set_error(const char* fmt, ...);
int check_launch(const char* what);

#define PTTA_CHECK(cond, ...)                                  \
    do {                                                       \
        if (!(cond)) {                                         \
            ::ptta::set_error(__VA_ARGS__);                    \
            return 1;                                          \
        }                                                      \
    } while (0)

#define PTTA_CUDA(call)                                                                   \
    do {                                                                                  \
        cudaError_t _e = (call);                                                          \
        if (_e != cudaSuccess) {                                                          \
            ::ptta::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e),     \
                              __FILE__, __LINE__);                                        \
            return 2;                                                                     \
        }                                                                                 \
    } while (0)

#define PTTA_TRY(expr)            \
    do {                          \
        int _rc = (expr);         \
        if (_rc) return _rc;      \
    } while (0)

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// ---- launches: programmatic dependent launch (PDL) ------------------------------------------------
// Every kernel of the library is launched with the programmatic-stream-serialization attribute and starts with PDL_SYNC():
// `griddepcontrol.wait` (all prerequisite grids complete, their memory visible) followed by `griddepcontrol.launch_dependents`
// (the NEXT kernel of the stream may be scheduled now: its blocks become resident as SM resources free up, run their
// set-up -- barrier init, TMEM allocation, descriptor prefetch -- and park in their own wait).  In a captured step this turns the
// kernel -> kernel edges into programmatic edges: launch latency and block scheduling of kernel k+1 overlap the tail of kernel
// k instead of following it.  The trigger comes AFTER the wait, so at most one successor is ever resident ahead of time and a
// kernel that reads data produced two launches back is still ordered behind it.  PTTA_PDL=0 in the environment disables it.
bool pdl_enabled();

template <typename... KArgs, typename... Args>
inline void launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 1 : 0;
    (void)cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);      // errors surface in check_launch()
}

// ---- device helpers -----------------------------------------------------------------------------
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#ifdef PTTA_STAMPS
// profiling build only (lib/libptta_b200_stamps.so, tools/graph_stamps.py): block 0 of every kernel records %globaltimer right
// after its dependency wait, so a REPLAYED graph yields the in-situ start time of each kernel (start-to-start = duration + gap)
__device__ unsigned long long g_stamps[8192];
__device__ unsigned int g_stamp_count;
__device__ __forceinline__ void pdl_stamp() {
    if ((blockIdx.x | blockIdx.y | blockIdx.z | threadIdx.x | threadIdx.y | threadIdx.z) == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        const unsigned int i = atomicAdd(&g_stamp_count, 1u);
        if (i < 8192u) g_stamps[i] = t;
    }
}
#define PDL_SYNC() do { ::ptta::pdl_wait(); ::ptta::pdl_trigger(); ::ptta::pdl_stamp(); } while (0)
#else
#define PDL_SYNC() do { ::ptta::pdl_wait(); ::ptta::pdl_trigger(); } while (0)
#endif

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ void ldmatrix_x4(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}

__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}

// D(16x8, fp32) += A(16x16, bf16, row) * B(16x8, bf16, col)
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ void cp_async16(uint32_t smem_addr, const void* gptr, bool valid) {
    int sz = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" :: "r"(smem_addr), "l"(gptr), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" :: "n"(N)); }

__device__ __forceinline__ uint32_t pack_bf162(float a, float b) {
    bf162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack_bf162(uint32_t u) {
    bf162 v = *reinterpret_cast<bf162*>(&u);
    return __bfloat1622float2(v);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// block-wide sum of `v` (blockDim.x multiple of 32, <= 1024); result valid in thread 0
__device__ __forceinline__ double block_sum_d(double v, double* sh /* >= 32 doubles */) {
    v = warp_sum_d(v);
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) sh[w] = v;
    __syncthreads();
    double r = 0.0;
    if (w == 0) {
        int nw = (blockDim.x + 31) >> 5;
        r = lane < nw ? sh[lane] : 0.0;
        r = warp_sum_d(r);
    }
    return r;
}

}  // namespace ptta
