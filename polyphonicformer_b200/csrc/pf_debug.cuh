// Optional in-kernel timeline (pf_debug_timeline): block (0,0,0) of an instrumented launch claims a 16-slot record in
// a caller-provided device buffer and stores %globaltimer samples into it.  One pointer per translation unit (no
// relocatable device code); pf_debug_timeline sets all of them.  Costs one predictable branch when switched off.
#pragma once
#include <cuda_runtime.h>

namespace pf {
static __device__ long long* g_dbg = nullptr;

__device__ __forceinline__ long long gtime() {
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// thread 0 of block (0,0,0): claim a record, tag it, stamp slot 0
__device__ __forceinline__ long long* dbg_claim(int tag) {
    if (!g_dbg || blockIdx.x || blockIdx.y || blockIdx.z || threadIdx.x) return nullptr;
    const unsigned long long slot = atomicAdd(reinterpret_cast<unsigned long long*>(g_dbg), 1ull);
    long long* d = g_dbg + 16 + slot * 16;
    d[0] = gtime();
    d[15] = tag;
    return d;
}
// thread 0 of EVERY block (per-CTA skew studies); slot 14 = linear block id
__device__ __forceinline__ long long* dbg_claim_all(int tag) {
    if (!g_dbg || threadIdx.x) return nullptr;
    const unsigned long long slot = atomicAdd(reinterpret_cast<unsigned long long*>(g_dbg), 1ull);
    long long* d = g_dbg + 16 + slot * 16;
    d[0] = gtime();
    d[14] = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
    d[15] = tag;
    return d;
}
#define DBG(i) do { if (dbg) dbg[i] = gtime(); } while (0)
}  // namespace pf
#define PF_DEFINE_DBG_SETTER(name)                                                        \
    namespace pf {                                                                        \
    int name(long long* p) { return (int)cudaMemcpyToSymbol(g_dbg, &p, sizeof(p)); }      \
    }
