// Device-side building blocks for the peer-memory collectives: system-scope signal primitives,
// the block-wise cross-GPU barrier, multimem (NVLS) accessors and the bucket tensor table.
#pragma once
#include "common.h"
#include "peer_mem.h"

namespace b200 {

constexpr int kMaxBucketTensors = 192;      // keeps the whole launch argument block under 4 KB
constexpr int kCommThreads = 512;

// One gradient (or parameter) participating in a flat bucket.
struct TensorSlot {
  void* ptr;            // device pointer (nullptr = no gradient this iteration)
  uint32_t numel;       // elements
  uint32_t off;         // element offset of the slot in the flat bucket (multiple of 8)
};

struct BucketTable {
  int count;
  uint32_t data_elems;  // end of the padded data region == start of the "used" flags
  uint32_t total_elems; // data + flags, multiple of 8
  uint32_t _pad;
  TensorSlot t[kMaxBucketTensors];
};

struct CommCtx {
  char* base;           // VA of rank 0's arena; rank r at base + r*stride
  char* mc_base;        // multicast VA of the arena (nullptr if NVLS unavailable)
  size_t stride;
  size_t pad_off;       // byte offset of this communicator's signal pads inside the arena
  int* error_word;      // host-mapped; non-zero after a barrier timeout
  unsigned long long timeout_ns;
  int rank, world;
};

#ifdef __CUDACC__

__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// Signal slot protocol (self-resetting, so it is CUDA-graph safe: no epoch argument to bake):
//   sender: spin CAS 0->1 with release.sys   receiver: spin CAS 1->0 with acquire.sys
__device__ __forceinline__ uint32_t cas_release_sys(uint32_t* addr, uint32_t cmp, uint32_t val) {
  uint32_t old;
  asm volatile("atom.global.release.sys.cas.b32 %0, [%1], %2, %3;" : "=r"(old) : "l"(addr), "r"(cmp), "r"(val) : "memory");
  return old;
}
__device__ __forceinline__ uint32_t cas_acquire_sys(uint32_t* addr, uint32_t cmp, uint32_t val) {
  uint32_t old;
  asm volatile("atom.global.acquire.sys.cas.b32 %0, [%1], %2, %3;" : "=r"(old) : "l"(addr), "r"(cmp), "r"(val) : "memory");
  return old;
}

__device__ __forceinline__ bool put_signal(uint32_t* addr, const CommCtx& c) {
  const unsigned long long t0 = globaltimer_ns();
  while (cas_release_sys(addr, 0u, 1u) != 0u) {
    if (globaltimer_ns() - t0 > c.timeout_ns) { *c.error_word = 100 + c.rank; return false; }
  }
  return true;
}
__device__ __forceinline__ bool wait_signal(uint32_t* addr, const CommCtx& c) {
  const unsigned long long t0 = globaltimer_ns();
  while (cas_acquire_sys(addr, 1u, 0u) != 1u) {
    if (globaltimer_ns() - t0 > c.timeout_ns) { *c.error_word = 200 + c.rank; return false; }
  }
  return true;
}

// Barrier between block `blockIdx.x` of every rank.  Pad layout: [block][sender] uint32.
// Writes made by any thread of this block before the call are visible to the peer blocks after
// their call returns (bar.sync -> release.sys ... acquire.sys -> bar.sync).
__device__ __forceinline__ void peer_block_barrier(const CommCtx& c) {
  __syncthreads();
  if ((int)threadIdx.x < c.world) {
    const int peer = threadIdx.x;
    uint32_t* remote = reinterpret_cast<uint32_t*>(c.base + (size_t)peer * c.stride + c.pad_off) +
                       (size_t)blockIdx.x * kMaxRanks + c.rank;
    uint32_t* mine = reinterpret_cast<uint32_t*>(c.base + (size_t)c.rank * c.stride + c.pad_off) +
                     (size_t)blockIdx.x * kMaxRanks + peer;
    if (put_signal(remote, c)) wait_signal(mine, c);
  }
  __syncthreads();
}

// ---- 16-byte accessors -----------------------------------------------------------------------
struct alignas(16) Vec16 { uint32_t w[4]; };

__device__ __forceinline__ Vec16 ld_sys(const void* p) {   // coherent load (peer memory)
  Vec16 v;
  asm volatile("ld.relaxed.sys.global.v4.b32 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.w[0]), "=r"(v.w[1]), "=r"(v.w[2]), "=r"(v.w[3]) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_sys(void* p, const Vec16& v) {
  asm volatile("st.relaxed.sys.global.v4.b32 [%0], {%1,%2,%3,%4};"
               :: "l"(p), "r"(v.w[0]), "r"(v.w[1]), "r"(v.w[2]), "r"(v.w[3]) : "memory");
}
__device__ __forceinline__ Vec16 ld_cg(const void* p) {    // local load, L1 bypass
  Vec16 v;
  asm volatile("ld.global.cg.v4.b32 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.w[0]), "=r"(v.w[1]), "=r"(v.w[2]), "=r"(v.w[3]) : "l"(p) : "memory");
  return v;
}

// NVLS: in-switch reduction of the same offset in every rank's arena, and broadcast store.
__device__ __forceinline__ Vec16 multimem_ld_reduce_bf16(const void* mc) {
  Vec16 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.bf16x2 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.w[0]), "=r"(v.w[1]), "=r"(v.w[2]), "=r"(v.w[3]) : "l"(mc) : "memory");
  return v;
}
__device__ __forceinline__ Vec16 multimem_ld_reduce_f32(const void* mc) {
  Vec16 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.w[0]), "=r"(v.w[1]), "=r"(v.w[2]), "=r"(v.w[3]) : "l"(mc) : "memory");
  return v;
}
__device__ __forceinline__ void multimem_st(void* mc, const Vec16& v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};"
               :: "l"(mc), "r"(v.w[0]), "r"(v.w[1]), "r"(v.w[2]), "r"(v.w[3]) : "memory");
}

// Wire-format traits: how many elements ride in one 16-byte vector.
template <typename WireT> struct Wire;
template <> struct Wire<__nv_bfloat16> {
  static constexpr int VE = 8;
  __device__ static void unpack(const Vec16& v, float* f) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(&v.w[i]);
      float2 t = __bfloat1622float2(h);
      f[2 * i] = t.x; f[2 * i + 1] = t.y;
    }
  }
  __device__ static Vec16 pack(const float* f) {
    Vec16 v;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
      v.w[i] = *reinterpret_cast<uint32_t*>(&h);
    }
    return v;
  }
  __device__ static Vec16 mc_reduce(const void* mc) { return multimem_ld_reduce_bf16(mc); }
};
template <> struct Wire<float> {
  static constexpr int VE = 4;
  __device__ static void unpack(const Vec16& v, float* f) {
#pragma unroll
    for (int i = 0; i < 4; ++i) f[i] = __uint_as_float(v.w[i]);
  }
  __device__ static Vec16 pack(const float* f) {
    Vec16 v;
#pragma unroll
    for (int i = 0; i < 4; ++i) v.w[i] = __float_as_uint(f[i]);
    return v;
  }
  __device__ static Vec16 mc_reduce(const void* mc) { return multimem_ld_reduce_f32(mc); }
};

// Which slot of the (shared-memory) offset array contains element e?  offs[count] == data_elems.
__device__ __forceinline__ int find_slot(const uint32_t* offs, int count, uint32_t e) {
  int lo = 0, hi = count - 1;
  while (lo < hi) {
    int mid = (lo + hi + 1) >> 1;
    if (offs[mid] <= e) lo = mid; else hi = mid - 1;
  }
  return lo;
}

#endif  // __CUDACC__

}  // namespace b200
