// Kernel skeletons for the scalar-field gates: one gate per thread, 256-bit global loads/stores
// (LDG.E.256 / STG.E.256 on sm_100a: a warp moves 1024 contiguous bytes per instruction on a
// planar vector), persistent grid-stride loops sized from the SM count.
//
// Every operand is (pointer, byte stride): planar vectors use stride 32, the reference's AoS
// ScalarShare image {share, mac} (share.rs:32-37) uses stride 64 with the mac plane at +32 bytes, so the
// same kernels serve device-resident planes and AoS chunks staged from host buffers.
#pragma once
#include <cuda_runtime.h>
#include "beaver.cuh"

namespace ark {

// A strided view of a vector of 32-byte field elements in global memory.
struct Vec {
  const char* p;
  uint32_t stride;  // bytes between consecutive elements (32 planar, 64 AoS)
};
struct MVec {
  char* p;
  uint32_t stride;
};

__device__ __forceinline__ void ld_fe(fe8& r, const Vec& v, size_t i) {
  const char* a = v.p + i * (size_t)v.stride;
  asm volatile("ld.global.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7])
               : "l"(a)
               : "memory");
}
__device__ __forceinline__ void st_fe(const MVec& v, size_t i, const fe8& r) {
  char* a = v.p + i * (size_t)v.stride;
  asm volatile("st.global.L1::no_allocate.v8.u32 [%8], {%0,%1,%2,%3,%4,%5,%6,%7};"
               :
               : "r"(r.v[0]), "r"(r.v[1]), "r"(r.v[2]), "r"(r.v[3]), "r"(r.v[4]), "r"(r.v[5]), "r"(r.v[6]), "r"(r.v[7]), "l"(a)
               : "memory");
}

// L2 eviction priorities for the Beaver kernels: operands that are read exactly once (triples, x, y) are marked evict-first and
// the masks K1 writes evict-last, so that what K2 reads back a moment later — the four d / e planes, 128 MB per 2^20-gate
// two-party step against a 126 MB L2 — is less likely to have been pushed out by the streams.
__device__ __forceinline__ void ld_fe_stream(fe8& r, const Vec& v, size_t i) {
  const char* a = v.p + i * (size_t)v.stride;
  asm volatile("ld.global.L1::no_allocate.L2::evict_first.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7])
               : "l"(a)
               : "memory");
}
__device__ __forceinline__ void st_fe_keep(const MVec& v, size_t i, const fe8& r) {
  char* a = v.p + i * (size_t)v.stride;
  asm volatile("st.global.L1::no_allocate.L2::evict_last.v8.u32 [%8], {%0,%1,%2,%3,%4,%5,%6,%7};"
               :
               : "r"(r.v[0]), "r"(r.v[1]), "r"(r.v[2]), "r"(r.v[3]), "r"(r.v[4]), "r"(r.v[5]), "r"(r.v[6]), "r"(r.v[7]), "l"(a)
               : "memory");
}

constexpr int kBlock = 256;

// Programmatic dependent launch (sm_90+): a kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization may be
// scheduled while its predecessor in the stream drains.  `pdl_prologue()` waits until every predecessor grid has completed and
// its writes are visible (nothing before it touches global memory), then lets OUR successor be scheduled.  Launched normally
// (<<<>>>), both instructions are no-ops.  Measured on the 4-kernel Beaver step: 213-221 -> 207-210 us (tools/_pdl.cu).
//
// `independent` (arkmpc_ctx_hint_independent): the host states that this launch is independent of the launches before it back
// to the most recent un-hinted one (it neither reads nor overwrites what they write or read — e.g. the other party's
// recombine in a mock run, or a gate on other operands).  The wait moves from the first instruction to the LAST: the grid's
// loads and arithmetic overlap the predecessor's drain — drain plus ramp-up cost ~8 us of a 70 us recombine launch otherwise
// (tools/_k2t.cu, profiles/r02b_k2_timeline.txt) — and the wait before exit keeps completion transitive: a later launch that
// waits on this grid also waits on everything before it.  An un-hinted launch triggers its successors only after its own
// wait has returned, so a hinted successor never runs ahead of work that is older than its group.
__device__ __forceinline__ void pdl_prologue(bool independent = false) {
  if (!independent) asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;");
}
__device__ __forceinline__ void pdl_epilogue(bool independent) {
  if (independent) asm volatile("griddepcontrol.wait;" ::: "memory");
}

// ---------------------------------------------------------------------------------------------
// Beaver phase 1: d_mine = x - a, e_mine = y - b on the share components.   192 B / gate.
// ---------------------------------------------------------------------------------------------
template <class F>
__global__ void __launch_bounds__(kBlock) beaver_mask_kernel(size_t n, Vec x, Vec y, Vec a, Vec b, MVec d, MVec e, int independent, int keep) {
  pdl_prologue(independent != 0);
  const size_t step = (size_t)gridDim.x * kBlock;
  for (size_t i = (size_t)blockIdx.x * kBlock + threadIdx.x; i < n; i += step) {
    fe8 xs, ys, as, bs, dm, em;
    ld_fe_stream(xs, x, i);
    ld_fe_stream(ys, y, i);
    ld_fe_stream(as, a, i);
    ld_fe_stream(bs, b, i);
    beaver_mask_elem<F>(dm, em, xs, ys, as, bs);
    if (keep) {
      st_fe_keep(d, i, dm);
      st_fe_keep(e, i, em);
    } else {
      st_fe(d, i, dm);
      st_fe(e, i, em);
    }
  }
  pdl_epilogue(independent != 0);
}

// ---------------------------------------------------------------------------------------------
// Beaver phase 2 (fused open-add + recombine + MAC update).   384 B / gate (+64 B with OPEN).
// ---------------------------------------------------------------------------------------------
struct RecombineArgs {
  Vec d_mine, e_mine, d_peer, e_peer;
  Vec a_s, a_m, b_s, b_m, c_s, c_m;
  MVec out_s, out_m, d_open, e_open;
  CTab key;  // constant-multiplier table of this party's MAC key share (ctab.hpp)
  int independent;  // see pdl_prologue
};

// Launch shape (measured on B200, tools/_k2v.cu, profiles/r01d_k2_launch_shapes.txt): the kernel is bound by the IMAD.WIDE
// pipe, not by occupancy, but a one-gate-per-thread grid (hardware block scheduling instead of a persistent grid-stride
// wave) with 3 resident blocks of 256 (78 registers) is 10 % faster than 2 persistent blocks per SM: 74.9 vs 82.7 us.
constexpr int kRecombineMinBlocks = 3;

// One gate per thread, no grid-stride loop: the launch covers n (gridDim.x < 2^31), and the straight-line body is 2.5 % faster
// than the same body inside a loop (tools/_k2t.cu vs tools/_k2v.cu).
template <class F, int PARTY, bool OPEN>
__global__ void __launch_bounds__(kBlock, kRecombineMinBlocks) beaver_recombine_kernel(size_t n, const __grid_constant__ RecombineArgs g) {
  pdl_prologue(g.independent != 0);
  const size_t i = (size_t)blockIdx.x * kBlock + threadIdx.x;
  if (i < n) {
    fe8 dm, em, dp, ep, as, am, bs, bm, cs, cm;
    ld_fe(dm, g.d_mine, i);
    ld_fe(dp, g.d_peer, i);
    ld_fe(em, g.e_mine, i);
    ld_fe(ep, g.e_peer, i);
    ld_fe_stream(bs, g.b_s, i);
    ld_fe_stream(as, g.a_s, i);
    ld_fe_stream(bm, g.b_m, i);
    ld_fe_stream(am, g.a_m, i);
    ld_fe_stream(cs, g.c_s, i);
    ld_fe_stream(cm, g.c_m, i);
    fe8 os, om, d, e;
    beaver_recombine_elem<F>(os, om, d, e, PARTY, g.key, dm, em, dp, ep, as, am, bs, bm, cs, cm);
    st_fe(g.out_s, i, os);
    st_fe(g.out_m, i, om);
    if (OPEN) {
      st_fe(g.d_open, i, d);
      st_fe(g.e_open, i, e);
    }
  }
  pdl_epilogue(g.independent != 0);
}

// ---------------------------------------------------------------------------------------------
// Beaver phase 2 with the batch_open all-gather fused into the stores (multi-GPU): the opened d and e of
// this rank's shard are written straight into rows [rank*n, (rank+1)*n) of EVERY rank's gathered planes through
// peer pointers (CUDA IPC mappings over NVLink / NVSwitch), so the transfer overlaps the arithmetic gate by gate and
// the separate d_open/e_open round trip through local HBM plus the NCCL launch disappear.
// ---------------------------------------------------------------------------------------------
constexpr int kMaxPeers = 8;
struct GatherArgs {
  char* d[kMaxPeers];  // rank k's gathered d plane, already offset to this rank's row block
  char* e[kMaxPeers];
  int world;
};

template <class F, int PARTY>
__global__ void __launch_bounds__(kBlock, kRecombineMinBlocks) beaver_recombine_gather_kernel(size_t n, const __grid_constant__ RecombineArgs g,
                                                                            const __grid_constant__ GatherArgs q) {
  pdl_prologue(g.independent != 0);
  const size_t step = (size_t)gridDim.x * kBlock;
  for (size_t i = (size_t)blockIdx.x * kBlock + threadIdx.x; i < n; i += step) {
    fe8 dm, em, dp, ep, as, am, bs, bm, cs, cm;
    ld_fe(dm, g.d_mine, i);
    ld_fe(dp, g.d_peer, i);
    ld_fe(em, g.e_mine, i);
    ld_fe(ep, g.e_peer, i);
    ld_fe_stream(bs, g.b_s, i);
    ld_fe_stream(as, g.a_s, i);
    ld_fe_stream(bm, g.b_m, i);
    ld_fe_stream(am, g.a_m, i);
    ld_fe_stream(cs, g.c_s, i);
    ld_fe_stream(cm, g.c_m, i);
    fe8 os, om, d, e;
    beaver_recombine_elem<F>(os, om, d, e, PARTY, g.key, dm, em, dp, ep, as, am, bs, bm, cs, cm);
    st_fe(g.out_s, i, os);
    st_fe(g.out_m, i, om);
#pragma unroll
    for (int k = 0; k < kMaxPeers; k++) {
      if (k < q.world) {
        st_fe(MVec{q.d[k], 32}, i, d);
        st_fe(MVec{q.e[k], 32}, i, e);
      }
    }
  }
  pdl_epilogue(g.independent != 0);
}

// ---------------------------------------------------------------------------------------------
// The same with NVSwitch MULTICAST stores: the gathered planes of all ranks are bound to one multicast object
// (arkmpc_mc_*, arkmpc_comm.cu) and each opened element leaves the SM ONCE (multimem.st, 2 x 16 B per plane); the switch
// replicates it into every rank's memory, this rank's own copy included.  Egress per rank drops from (world-1) x 64 B
// to 64 B per gate, so the gather stops being bound by the rank's NVLink egress.
// ---------------------------------------------------------------------------------------------
struct McGatherArgs {
  char* d;  // multicast address of the gathered d plane, already offset to this rank's row block
  char* e;
};

__device__ __forceinline__ void st_fe_multicast(char* base, size_t i, const fe8& r) {
  char* a = base + i * 32;
  asm volatile("multimem.st.weak.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(a), "f"(__uint_as_float(r.v[0])), "f"(__uint_as_float(r.v[1])),
               "f"(__uint_as_float(r.v[2])), "f"(__uint_as_float(r.v[3]))
               : "memory");
  asm volatile("multimem.st.weak.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(a + 16), "f"(__uint_as_float(r.v[4])), "f"(__uint_as_float(r.v[5])),
               "f"(__uint_as_float(r.v[6])), "f"(__uint_as_float(r.v[7]))
               : "memory");
}

template <class F, int PARTY>
__global__ void __launch_bounds__(kBlock, kRecombineMinBlocks) beaver_recombine_gather_mc_kernel(size_t n, const __grid_constant__ RecombineArgs g,
                                                                                               const __grid_constant__ McGatherArgs q) {
  pdl_prologue(g.independent != 0);
  const size_t step = (size_t)gridDim.x * kBlock;
  for (size_t i = (size_t)blockIdx.x * kBlock + threadIdx.x; i < n; i += step) {
    fe8 dm, em, dp, ep, as, am, bs, bm, cs, cm;
    ld_fe(dm, g.d_mine, i);
    ld_fe(dp, g.d_peer, i);
    ld_fe(em, g.e_mine, i);
    ld_fe(ep, g.e_peer, i);
    ld_fe_stream(bs, g.b_s, i);
    ld_fe_stream(as, g.a_s, i);
    ld_fe_stream(bm, g.b_m, i);
    ld_fe_stream(am, g.a_m, i);
    ld_fe_stream(cs, g.c_s, i);
    ld_fe_stream(cm, g.c_m, i);
    fe8 os, om, d, e;
    beaver_recombine_elem<F>(os, om, d, e, PARTY, g.key, dm, em, dp, ep, as, am, bs, bm, cs, cm);
    st_fe(g.out_s, i, os);
    st_fe(g.out_m, i, om);
    st_fe_multicast(q.d, i, d);
    st_fe_multicast(q.e, i, e);
  }
  pdl_epilogue(g.independent != 0);
}

// plain all-gather by multicast (no arithmetic): rows of this rank's local plane -> every rank's gathered plane
static __global__ void __launch_bounds__(kBlock) multicast_rows_kernel(size_t n, Vec local, char* mc_rows) {
  const size_t step = (size_t)gridDim.x * kBlock;
  for (size_t i = (size_t)blockIdx.x * kBlock + threadIdx.x; i < n; i += step) {
    fe8 v;
    ld_fe(v, local, i);
    st_fe_multicast(mc_rows, i, v);
  }
}

// ---------------------------------------------------------------------------------------------
// Beaver phase 2, TMA-staged variant for planar (stride-32) operands.
//
// The LDG kernel above keeps only ~16 warps x 320 B in flight per SM and every warp stalls on its ten loads before
// its ~850 arithmetic instructions, so neither HBM nor the IMAD pipe saturates (ncu r01a: DRAM 56 %, fma pipe 41 %).
// Here every warp owns one 10 KiB shared-memory slot (ten planes x 32 elements x 32 B) and one mbarrier.  Per tile
// of 32 gates the warp (1) waits for the slot, (2) copies its operands into registers (2 x LDS.128 per operand),
// (3) lane 0 immediately re-arms the barrier and issues the ten 1 KiB bulk copies (cp.async.bulk -> SASS UBLKCP)
// of the warp's NEXT tile into the same slot, (4) all lanes run the fused recombination and store with STG.256.
// The copy of tile k+1 therefore overlaps the whole arithmetic of tile k with zero register cost, and warps
// drift apart so HBM demand is smooth.  2 CTAs x 8 warps x 10 KiB = 160 KiB of the SM's 227 KiB.
// ---------------------------------------------------------------------------------------------
constexpr int kTmaPlanes = 10;
constexpr int kTmaTile = 32;                                   // gates per warp tile
constexpr int kTmaPlaneBytes = kTmaTile * 32;                  // 1 KiB
constexpr int kTmaSlotBytes = kTmaPlanes * kTmaPlaneBytes;     // 10 KiB per warp
constexpr int kTmaWarps = kBlock / 32;
constexpr int kTmaSmemBytes = kTmaWarps * kTmaSlotBytes + kTmaWarps * 8 + kBlock * 4;  // slots | mbarriers | per-lane sink words

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void lds_fe(fe8& r, uint32_t addr) {
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]) : "r"(addr) : "memory");
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7]) : "r"(addr + 16) : "memory");
}

template <class F, int PARTY, bool OPEN>
__global__ void __launch_bounds__(kBlock, 2) beaver_recombine_tma_kernel(size_t n, const __grid_constant__ RecombineArgs g) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t slot = smem_u32(smem) + warp * kTmaSlotBytes;
  const uint32_t bar = smem_u32(smem) + kTmaWarps * kTmaSlotBytes + warp * 8;
  const uint32_t sink = smem_u32(smem) + kTmaWarps * kTmaSlotBytes + kTmaWarps * 8 + threadIdx.x * 4;
  const size_t tiles = (n + kTmaTile - 1) / kTmaTile;
  const size_t stride = (size_t)gridDim.x * kTmaWarps;
  size_t t = (size_t)blockIdx.x * kTmaWarps + warp;

  // order in the slot: d_mine d_peer e_mine e_peer b_s a_s b_m a_m c_s c_m
  auto issue = [&](size_t tile) {
    const size_t first = tile * kTmaTile;
    const uint32_t bytes = (uint32_t)((n - first < (size_t)kTmaTile ? n - first : (size_t)kTmaTile) * 32);
    const size_t off = first * 32;
    mbar_expect_tx(bar, bytes * kTmaPlanes);
    bulk_g2s(slot + 0 * kTmaPlaneBytes, g.d_mine.p + off, bytes, bar);
    bulk_g2s(slot + 1 * kTmaPlaneBytes, g.d_peer.p + off, bytes, bar);
    bulk_g2s(slot + 2 * kTmaPlaneBytes, g.e_mine.p + off, bytes, bar);
    bulk_g2s(slot + 3 * kTmaPlaneBytes, g.e_peer.p + off, bytes, bar);
    bulk_g2s(slot + 4 * kTmaPlaneBytes, g.b_s.p + off, bytes, bar);
    bulk_g2s(slot + 5 * kTmaPlaneBytes, g.a_s.p + off, bytes, bar);
    bulk_g2s(slot + 6 * kTmaPlaneBytes, g.b_m.p + off, bytes, bar);
    bulk_g2s(slot + 7 * kTmaPlaneBytes, g.a_m.p + off, bytes, bar);
    bulk_g2s(slot + 8 * kTmaPlaneBytes, g.c_s.p + off, bytes, bar);
    bulk_g2s(slot + 9 * kTmaPlaneBytes, g.c_m.p + off, bytes, bar);
  };

  if (lane == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if (t < tiles) issue(t);
  }
  __syncwarp();

  uint32_t parity = 0;
  for (; t < tiles; t += stride) {
    mbar_wait(bar, parity);
    parity ^= 1;
    const uint32_t mine = slot + lane * 32;
    fe8 dm, dp, em, ep, bs, as, bm, am, cs, cm;
    lds_fe(dm, mine + 0 * kTmaPlaneBytes);
    lds_fe(dp, mine + 1 * kTmaPlaneBytes);
    lds_fe(em, mine + 2 * kTmaPlaneBytes);
    lds_fe(ep, mine + 3 * kTmaPlaneBytes);
    lds_fe(bs, mine + 4 * kTmaPlaneBytes);
    lds_fe(as, mine + 5 * kTmaPlaneBytes);
    lds_fe(bm, mine + 6 * kTmaPlaneBytes);
    lds_fe(am, mine + 7 * kTmaPlaneBytes);
    lds_fe(cs, mine + 8 * kTmaPlaneBytes);
    lds_fe(cm, mine + 9 * kTmaPlaneBytes);
    // Every LDS must have RETURNED before the slot is handed back to the async proxy: consume one register of
    // each load (the scoreboard wait happens at first use), then converge the warp.
    uint32_t seen = (dm.v[0] | dm.v[4]) ^ (dp.v[0] | dp.v[4]) ^ (em.v[0] | em.v[4]) ^ (ep.v[0] | ep.v[4]) ^ (bs.v[0] | bs.v[4]) ^
                    (as.v[0] | as.v[4]) ^ (bm.v[0] | bm.v[4]) ^ (am.v[0] | am.v[4]) ^ (cs.v[0] | cs.v[4]) ^ (cm.v[0] | cm.v[4]);
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(sink), "r"(seen) : "memory");  // a side effect ptxas cannot drop
    __syncwarp();
    const size_t next = t + stride;
    if (lane == 0 && next < tiles) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      issue(next);
    }
    const size_t i = t * kTmaTile + lane;
    if (i < n) {
      fe8 os, om, d, e;
      beaver_recombine_elem<F>(os, om, d, e, PARTY, g.key, dm, em, dp, ep, as, am, bs, bm, cs, cm);
      st_fe(g.out_s, i, os);
      st_fe(g.out_m, i, om);
      if (OPEN) {
        st_fe(g.d_open, i, d);
        st_fe(g.e_open, i, e);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Generic element-wise gates
// ---------------------------------------------------------------------------------------------
enum class Bin { Add, Sub, Mul };

template <class F, Bin OP>
__global__ void __launch_bounds__(kBlock) fr_binary_kernel(size_t n, Vec a, Vec b, MVec out) {
  const size_t step = (size_t)gridDim.x * kBlock;
  for (size_t i = (size_t)blockIdx.x * kBlock + threadIdx.x; i < n; i += step) {
    fe8 x, y, r;
    ld_fe(x, a, i);
    ld_fe(y, b, i);
    if (OP == Bin::Add) Fp<F>::add(r, x, y);
    if (OP == Bin::Sub) Fp<F>::sub(r, x, y);
    if (OP == Bin::Mul) Fp<F>::mul(r, x, y);
    st_fe(out, i, r);
  }
}

// two planes at once: (a_s op b_s, a_m op b_m) -- share add / sub
template <class F, Bin OP>
__global__ void __launch_bounds__(kBlock) fr_share_binary_kernel(size_t n, Vec a_s, Vec a_m, Vec b_s, Vec b_m, MVec out_s, MVec out_m) {
  const size_t step = (size_t)gridDim.x * kBlock;
  for (size_t i = (size_t)blockIdx.x * kBlock + threadIdx.x; i < n; i += step) {
    fe8 x, y, r;
    ld_fe(x, a_s, i);
    ld_fe(y, b_s, i);
    if (OP == Bin::Add) Fp<F>::add(r, x, y); else Fp<F>::sub(r, x, y);
    st_fe(out_s, i, r);
    ld_fe(x, a_m, i);
    ld_fe(y, b_m, i);
    if (OP == Bin::Add) Fp<F>::add(r, x, y); else Fp<F>::sub(r, x, y);
    st_fe(out_m, i, r);
  }
}

template <class F>
__global__ void __launch_bounds__(kBlock) fr_neg_kernel(size_t n, Vec a, MVec out) {
  const size_t step = (size_t)gridDim.x * kBlock;
  for (size_t i = (size_t)blockIdx.x * kBlock + threadIdx.x; i < n; i += step) {
    fe8 x, r;
    ld_fe(x, a, i);
    Fp<F>::neg(r, x);
    st_fe(out, i, r);
  }
}

// out = a * s for one broadcast scalar s given as its constant-multiplier table (also to_mont with s = R^2, from_mont with s = 1)
template <class F>
__global__ void __launch_bounds__(kBlock) fr_scale_kernel(size_t n, Vec a, const __grid_constant__ CTab s, MVec out) {
  const size_t step = (size_t)gridDim.x * kBlock;
  for (size_t i = (size_t)blockIdx.x * kBlock + threadIdx.x; i < n; i += step) {
    fe8 x, r;
    ld_fe(x, a, i);
    Fp<F>::mul_ctab(r, s, x);
    st_fe(out, i, r);
  }
}

// (share*v, mac*v)
template <class F>
__global__ void __launch_bounds__(kBlock) fr_share_mul_public_kernel(size_t n, Vec a_s, Vec a_m, Vec v, MVec out_s, MVec out_m) {
  const size_t step = (size_t)gridDim.x * kBlock;
  for (size_t i = (size_t)blockIdx.x * kBlock + threadIdx.x; i < n; i += step) {
    fe8 s, m, w, r;
    ld_fe(w, v, i);
    ld_fe(s, a_s, i);
    Fp<F>::mul(r, s, w);
    st_fe(out_s, i, r);
    ld_fe(m, a_m, i);
    Fp<F>::mul(r, m, w);
    st_fe(out_m, i, r);
  }
}

template <class F, int PARTY, bool SUB>
__global__ void __launch_bounds__(kBlock) fr_share_add_public_kernel(size_t n, Vec a_s, Vec a_m, Vec v, const __grid_constant__ CTab key, MVec out_s, MVec out_m) {
  const size_t step = (size_t)gridDim.x * kBlock;
  for (size_t i = (size_t)blockIdx.x * kBlock + threadIdx.x; i < n; i += step) {
    fe8 s, m, w, os, om;
    ld_fe(s, a_s, i);
    ld_fe(m, a_m, i);
    ld_fe(w, v, i);
    if (SUB) share_sub_public_elem<F>(os, om, PARTY, key, s, m, w); else share_add_public_elem<F>(os, om, PARTY, key, s, m, w);
    st_fe(out_s, i, os);
    st_fe(out_m, i, om);
  }
}

template <class F>
__global__ void __launch_bounds__(kBlock) fr_mac_check_kernel(size_t n, Vec opened, Vec mac, const __grid_constant__ CTab key, MVec out) {
  const size_t step = (size_t)gridDim.x * kBlock;
  for (size_t i = (size_t)blockIdx.x * kBlock + threadIdx.x; i < n; i += step) {
    fe8 o, m, r;
    ld_fe(o, opened, i);
    ld_fe(m, mac, i);
    mac_check_elem<F>(r, key, o, m);
    st_fe(out, i, r);
  }
}

// flag (initialised to 1 by the host) is cleared if any mine[i] + peer[i] != 0
template <class F>
__global__ void __launch_bounds__(kBlock) fr_sum_is_zero_kernel(size_t n, Vec mine, Vec peer, int* flag) {
  const size_t step = (size_t)gridDim.x * kBlock;
  bool ok = true;
  for (size_t i = (size_t)blockIdx.x * kBlock + threadIdx.x; i < n; i += step) {
    fe8 x, y, r;
    ld_fe(x, mine, i);
    ld_fe(y, peer, i);
    Fp<F>::add(r, x, y);
    ok = ok && Fp<F>::is_zero(r);
  }
  if (!__all_sync(0xffffffffu, ok) && (threadIdx.x & 31) == 0) atomicAnd(flag, 0);
}

// flag (initialised to 1 by the host) is cleared if any a[i] >= p: what arkworks' deserialisation rejects (scalar.rs:187-202)
template <class F>
__global__ void __launch_bounds__(kBlock) fr_validate_kernel(size_t n, Vec a, int* flag) {
  const size_t step = (size_t)gridDim.x * kBlock;
  bool ok = true;
  for (size_t i = (size_t)blockIdx.x * kBlock + threadIdx.x; i < n; i += step) {
    fe8 x;
    ld_fe(x, a, i);
    ok = ok && Fp<F>::is_canonical(x);
  }
  if (!__all_sync(0xffffffffu, ok) && (threadIdx.x & 31) == 0) atomicAnd(flag, 0);
}

// canonical integer value as 32 big-endian bytes (scalar.rs:118-127); `one` = table of the integer 1 (leaves Montgomery form)
template <class F>
__global__ void __launch_bounds__(kBlock) fr_to_bytes_be_kernel(size_t n, Vec a, const __grid_constant__ CTab one, MVec out) {
  const size_t step = (size_t)gridDim.x * kBlock;
  for (size_t i = (size_t)blockIdx.x * kBlock + threadIdx.x; i < n; i += step) {
    fe8 x, r, o;
    ld_fe(x, a, i);
    Fp<F>::mul_ctab(r, one, x);
    for (int j = 0; j < 8; j++) o.v[j] = __byte_perm(r.v[7 - j], 0, 0x0123);
    st_fe(out, i, o);
  }
}

// ---------------------------------------------------------------------------------------------
// Sum (share.rs:104-111): thread-serial over a grid-stride slice, then warp-shuffle tree, then
// one partial per block; a second single-block launch folds the partials.
// ---------------------------------------------------------------------------------------------
template <class F>
__device__ __forceinline__ void warp_sum(fe8& acc) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    fe8 o, r;
#pragma unroll
    for (int j = 0; j < 8; j++) o.v[j] = __shfl_down_sync(0xffffffffu, acc.v[j], off);
    Fp<F>::add(r, acc, o);
    acc = r;
  }
}

template <class F>
__device__ __forceinline__ void block_sum_store(fe8 acc, fe8* smem /*[8]*/, MVec out, size_t out_index) {
  warp_sum<F>(acc);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) smem[warp] = acc;
  __syncthreads();
  if (warp == 0) {
    fe8 v;
    if (lane < kBlock / 32) v = smem[lane]; else Fp<F>::set_zero(v);
    warp_sum<F>(v);
    if (lane == 0) st_fe(out, out_index, v);
  }
  __syncthreads();
}

// NPLANES planes summed independently in one pass (2 for share+mac)
template <class F, int NPLANES>
__global__ void __launch_bounds__(kBlock) fr_sum_kernel(size_t n, Vec a0, Vec a1, MVec out0, MVec out1, bool per_block) {
  __shared__ fe8 smem[kBlock / 32];
  const size_t step = (size_t)gridDim.x * kBlock;
  fe8 acc0, acc1;
  Fp<F>::set_zero(acc0);
  Fp<F>::set_zero(acc1);
  for (size_t i = (size_t)blockIdx.x * kBlock + threadIdx.x; i < n; i += step) {
    fe8 x, r;
    ld_fe(x, a0, i);
    Fp<F>::add(r, acc0, x);
    acc0 = r;
    if (NPLANES == 2) {
      ld_fe(x, a1, i);
      Fp<F>::add(r, acc1, x);
      acc1 = r;
    }
  }
  const size_t oi = per_block ? blockIdx.x : 0;
  block_sum_store<F>(acc0, smem, out0, oi);
  if (NPLANES == 2) block_sum_store<F>(acc1, smem, out1, oi);
}

// ---------------------------------------------------------------------------------------------
// Beaver phase 2 fused with the Sum that follows it in an inner product (circuits.rs:22-50: sum_i [a_i][b_i]): the products
// are never written; each block leaves one partial (share, mac) and fr_sum_kernel folds the partials.  Modular addition is
// associative and commutative on canonical residues, so the result is bit-identical to batch_mul followed by sum().
// ---------------------------------------------------------------------------------------------
// Launch shape (profiles/r01k_recombine_sum_ab.txt): kSumGatesPerThread gates per thread (grid-stride, coalesced) and one partial per WARP — no shared memory, no
// barrier; with one gate per thread and a block-level tree the reduction cost more than the separate Sum launches it replaces.
constexpr int kSumGatesPerThread = 8;

template <class F, int PARTY>
__global__ void __launch_bounds__(kBlock, kRecombineMinBlocks) beaver_recombine_sum_kernel(size_t n, const __grid_constant__ RecombineArgs g,
                                                                                         MVec part_s, MVec part_m) {
  pdl_prologue(g.independent != 0);
  const size_t step = (size_t)gridDim.x * kBlock;
  fe8 acc_s, acc_m;
  Fp<F>::set_zero(acc_s);
  Fp<F>::set_zero(acc_m);
#pragma unroll 1
  for (size_t i = (size_t)blockIdx.x * kBlock + threadIdx.x; i < n; i += step) {
    fe8 dm, em, dp, ep, as, am, bs, bm, cs, cm;
    ld_fe(dm, g.d_mine, i);
    ld_fe(dp, g.d_peer, i);
    ld_fe(em, g.e_mine, i);
    ld_fe(ep, g.e_peer, i);
    ld_fe_stream(bs, g.b_s, i);
    ld_fe_stream(as, g.a_s, i);
    ld_fe_stream(bm, g.b_m, i);
    ld_fe_stream(am, g.a_m, i);
    ld_fe_stream(cs, g.c_s, i);
    ld_fe_stream(cm, g.c_m, i);
    fe8 os, om, d, e, r;
    beaver_recombine_elem<F>(os, om, d, e, PARTY, g.key, dm, em, dp, ep, as, am, bs, bm, cs, cm);
    Fp<F>::add(r, acc_s, os);
    acc_s = r;
    Fp<F>::add(r, acc_m, om);
    acc_m = r;
  }
  warp_sum<F>(acc_s);
  warp_sum<F>(acc_m);
  if ((threadIdx.x & 31) == 0) {
    const size_t w = (size_t)blockIdx.x * (kBlock / 32) + (threadIdx.x >> 5);
    st_fe(part_s, w, acc_s);
    st_fe(part_m, w, acc_m);
  }
  pdl_epilogue(g.independent != 0);
}

// ---------------------------------------------------------------------------------------------
// Layout conversion and synthetic data
// ---------------------------------------------------------------------------------------------
static __global__ void __launch_bounds__(kBlock) copy_planes_kernel(size_t n, Vec in_s, Vec in_m, MVec out_s, MVec out_m) {
  const size_t step = (size_t)gridDim.x * kBlock;
  for (size_t i = (size_t)blockIdx.x * kBlock + threadIdx.x; i < n; i += step) {
    fe8 s, m;
    ld_fe(s, in_s, i);
    ld_fe(m, in_m, i);
    st_fe(out_s, i, s);
    st_fe(out_m, i, m);
  }
}

__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  uint64_t z = x;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

// Uniform element of [0,p) from (seed, index), rejection-sampled; bit-identical to
// oracle/pyoracle.synth_element and oracle/ark_oracle.c:synth_one.
template <class F>
__global__ void __launch_bounds__(kBlock) fr_random_kernel(size_t n, uint64_t seed, uint64_t first, MVec out) {
  const size_t step = (size_t)gridDim.x * kBlock;
  const uint64_t top_mask = (1ull << (F::kBits - 192)) - 1;
  for (size_t i = (size_t)blockIdx.x * kBlock + threadIdx.x; i < n; i += step) {
    const uint64_t index = first + i;
    fe8 v;
    for (uint64_t t = 0;; t++) {
      uint64_t l[4];
#pragma unroll
      for (int j = 0; j < 4; j++) l[j] = splitmix64((seed ^ splitmix64(index * 4 + (uint64_t)j)) + t * 0xD1342543DE82EF95ull);
      l[3] &= top_mask;
#pragma unroll
      for (int j = 0; j < 4; j++) {
        v.v[2 * j] = (uint32_t)l[j];
        v.v[2 * j + 1] = (uint32_t)(l[j] >> 32);
      }
      if (Fp<F>::is_canonical(v)) break;
    }
    st_fe(out, i, v);
  }
}

}  // namespace ark
