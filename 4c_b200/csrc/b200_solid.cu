// b200_solid.cu -- lib4c_b200_solid.so: context, set-up kernels, evaluate drivers and the C ABI
// declared in include/b200_solid.h.  sm_100a only; FP64 throughout; no CPU fallback.
#include <cub/cub.cuh>
#include <cuda_runtime.h>
#ifndef B200S_L2_HINTS
#define B200S_L2_HINTS 0  // experiment switch: bit 0 = evict-first loads of the scatter map, bit 1 = evict-last REDs
#endif
#include <dlfcn.h>
#include <nvtx3/nvToolsExt.h>  // header-only: ranges are no-ops unless a profiler injects itself
#include <sys/mman.h>
#include <sys/syscall.h>
#include <unistd.h>

#include <cctype>
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <condition_variable>
#include <deque>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>
#if defined(__x86_64__)
#include <emmintrin.h>
#endif

// Translation-unit partitioning: the ~100 heavy kernel instantiations make one nvcc run take ~3 minutes, so
// __graft_entry__.build_library compiles this file five times in parallel with -DB200S_PART=0..4 (core + ABI,
// hex8, hex8 EAS, tet10, hex27) and links the objects.  Without the macro everything is one translation unit.
// (nvcc --split-compile was measured: 3x faster to build but 8-25 % slower kernels.)
#ifndef B200S_TET10_SPLIT_PHASE1
#define B200S_TET10_SPLIT_PHASE1 0  // 1: operator M formed after the X/u refill (see tet10_kernel.cuh phase1a / phase1b)
#endif
#define B200S_PART_CORE 0
#define B200S_PART_HEX8 1
#define B200S_PART_EAS 2
#define B200S_PART_TET10 3
#define B200S_PART_HEX27 4
#ifndef B200S_PART
#define B200S_HAS(p) 1
#else
#define B200S_HAS(p) (B200S_PART == (p))
#endif

#include "../../include/b200_solid.h"
#include "hex8_kernel.cuh"
#include "hex8_fbar_kernel.cuh"
#include "hex8_eas_kernel.cuh"
#include "hex8_eas_mma_kernel.cuh"
#include "tet10_kernel.cuh"
#include "hex27_kernel.cuh"
#if B200S_HAS(B200S_PART_CORE)
#include "post_kernels.cuh"
#include "generic_kernel.cuh"
#include "linalg_kernels.cuh"
#include "neumann_kernels.cuh"
#endif
#include "shape_tables.hpp"

namespace b200s
{
  // ------------------------------------------------------------------------------------------------
  // constant tables (per cell type; filled at set_mesh)
  // ------------------------------------------------------------------------------------------------
  struct Hex8Const
  {
    double dN_gp[8 * 24];
    double dN_node[8 * 24];
    double w[8];
    double xi_gp[8 * 3];  // parameter coordinates of the Gauss points (EAS shape functions)
  };
  static __constant__ Hex8Const c_hex8;  // one copy per translation unit that launches hex8 kernels
  struct Tet10Const
  {
    double dN_gp[4 * 30];
    double dN_node[10 * 30];
    double w[4];
  };
  static __constant__ Tet10Const c_tet10;
  struct Hex27Const
  {
    double dN_gp[27 * 81];
    double w[28];
    hex27::IntTables it;
  };
  static __device__ Hex27Const c_hex27;  // global memory (17.8 KB): copied to shared memory once per CTA

  // ------------------------------------------------------------------------------------------------
  // hex8 evaluate + assemble kernel (see hex8_kernel.cuh for the phase structure)
  // ------------------------------------------------------------------------------------------------
  constexpr int HEX8_WARPS = 4;  // warps per CTA; 2 CTAs/SM = 8 warps = 32 elements in flight / SM
  constexpr int HEX8_FBAR_WARPS = 7;  // F-bar variant: one CTA of 7 warps per SM (shared memory bound)
  constexpr size_t hex8_warp_bytes = (size_t)hex8::WARP_BYTES;
  constexpr size_t hex8_fbar_warp_bytes = (size_t)hex8fbar::WARP_BYTES;
  static_assert(hex8_fbar_warp_bytes % 16 == 0, "16-byte aligned carve-up");
  constexpr size_t hex8_table_bytes = (size_t)(2 * hex8::TAB_SIZE + 8) * 8;
  constexpr size_t hex8_smem_bytes = hex8_table_bytes + HEX8_WARPS * hex8_warp_bytes;
  constexpr size_t hex8_fbar_smem_bytes = hex8_table_bytes + HEX8_FBAR_WARPS * hex8_fbar_warp_bytes;
  static_assert(hex8_fbar_smem_bytes <= 232448, "fits the 227 KB of dynamic shared memory per CTA");
  static_assert(hex8_warp_bytes % 16 == 0 && hex8_table_bytes % 16 == 0, "16-byte aligned carve-up");

  __device__ __forceinline__ void atomic_min_double(double* addr, double v)
  {
    unsigned long long* a = reinterpret_cast<unsigned long long*>(addr);
    unsigned long long old = *a;
    while (__longlong_as_double((long long)old) > v)
    {
      const unsigned long long assumed = old;
      old = atomicCAS(a, assumed, (unsigned long long)__double_as_longlong(v));
      if (old == assumed) break;
    }
  }

  // cp.async (LDGSTS): global -> shared without a register round trip; completion per thread via
  // wait_group, visibility to the other lanes through the following __syncwarp()
  __device__ __forceinline__ void cp_async8(void* smem, const void* gmem)
  {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(
                     (unsigned)__cvta_generic_to_shared(smem)),
                 "l"(gmem)
                 : "memory");
  }
  __device__ __forceinline__ void cp_async16(void* smem, const void* gmem)
  {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(
                     (unsigned)__cvta_generic_to_shared(smem)),
                 "l"(gmem)
                 : "memory");
  }
  // streaming variant for data read exactly once per evaluate (the element -> CSR-slot map): L2 evict-first, so that
  // the map does not push partially accumulated CSR rows out of L2 (B200S_L2_HINTS bit 0; A/B in profiles/r02w_summary.md)
  __device__ __forceinline__ void cp_async16_stream(void* smem, const void* gmem)
  {
#if (B200S_L2_HINTS & 1)
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(
                     (unsigned)__cvta_generic_to_shared(smem)),
                 "l"(gmem), "l"(pol)
                 : "memory");
#else
    cp_async16(smem, gmem);
#endif
  }
  __device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
  // all but the most recently committed group are complete
  __device__ __forceinline__ void cp_async_wait_but_last() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }
  __device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

  // TMA bulk copy (cp.async.bulk, SASS UBLKCP) of a contiguous, 16-byte aligned block global -> shared, completion
  // signalled on an mbarrier (complete_tx); issued by ONE lane, awaited by all lanes of the warp
  __device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count)
  {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __device__ __forceinline__ void tma_bulk_load(void* smem, const void* gmem, unsigned bytes, uint64_t* bar)
  {
    const unsigned b = (unsigned)__cvta_generic_to_shared(bar);
    // earlier generic-proxy reads of the destination (the previous tile's scatter) are ordered before the
    // async-proxy write
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     (unsigned)__cvta_generic_to_shared(smem)),
                 "l"(gmem), "r"(bytes), "r"(b)
                 : "memory");
  }
  __device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity)
  {
    const unsigned b = (unsigned)__cvta_generic_to_shared(bar);
    unsigned done = 0;
    while (!done)
      asm volatile(
          "{\n\t.reg .pred p;\n\t"
          "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
          "selp.u32 %0, 1, 0, p;\n\t}"
          : "=r"(done)
          : "r"(b), "r"(parity)
          : "memory");
  }

  namespace hex8
  {
    // node id handled by this lane (element e0 + lane/8, local node lane%8) or -1 beyond the mesh
    __device__ __forceinline__ int load_node(const EvalArgs& A, int lane, int64_t tile, int64_t n_tiles)
    {
      if (tile >= n_tiles) return -1;
      const int64_t e = tile * EPW + (lane >> 3);
      return e < A.n_ele ? __ldg(A.conn + e * NN + (lane & 7)) : -1;
    }
    // X and u of this lane's node -> S.xu (6 x 8 bytes)
    __device__ __forceinline__ void issue_xu(const EvalArgs& A, const WarpSmem& S, int lane, int node)
    {
      if (node < 0)
      {
        gather_pad(S, lane);
        return;
      }
      double* xu = S.xu + (lane >> 3) * XU_ESTRIDE + 3 * (lane & 7);
      const double* gx = A.X + 3 * (int64_t)node;
      const double* gu = A.u + 3 * (int64_t)node;
#pragma unroll
      for (int d = 0; d < 3; ++d)
      {
        cp_async8(xu + d, gx + d);
        cp_async8(xu + 24 + d, gu + d);
      }
    }
    // CSR row info of this lane's node and the tile's 4 x 64 scatter-map entries (2 x 16 bytes/lane); with TMA the
    // tile's map (1 KB, contiguous) is ONE bulk copy issued by lane 0 and awaited on the warp's mbarrier
    template <bool TMA = false>
    __device__ __forceinline__ void issue_rowpos(
        const EvalArgs& A, const WarpSmem& S, int lane, int node, int64_t tile, uint64_t* bar = nullptr)
    {
      if (node >= 0)
        cp_async16(S.nrow + lane, A.node_row + node);
      else
        S.nrow[lane].rowstart = -1, S.nrow[lane].rowlen = 0, S.nrow[lane].frow = -1;
      const int64_t e0 = tile * EPW;
      if (TMA)
      {
        if (lane == 0)
        {
          const int64_t live = A.n_ele - e0 < EPW ? A.n_ele - e0 : EPW;
          tma_bulk_load(S.pos, A.pos + e0 * (NN * NN), (unsigned)(live * NN * NN * 4), bar);
        }
        return;
      }
#pragma unroll
      for (int h = 0; h < 2; ++h)
      {
        const int chunk = lane + 32 * h;  // 16-byte chunk of the tile's 1 KB map; 16 chunks / element
        if (e0 + (chunk >> 4) < A.n_ele)
          cp_async16_stream(S.pos + 4 * chunk, A.pos + e0 * (NN * NN) + 4 * chunk);
      }
    }
  }  // namespace hex8

#if B200S_HAS(B200S_PART_HEX8)
  // TECH 0: displacement based (4 warps, 2 CTAs/SM); TECH 1: F-bar (7 warps, 1 CTA/SM, more shared
  // memory per warp and the full non-symmetric block row per lane, hex8_fbar_kernel.cuh);
  // TECH 2: displacement based with MULF prestress (TECH 0 with the history-based deformation gradient);
  // TECH 3: displacement based with linear kinematics (KINEM linear)
  // TMA: the tile's scatter map is staged with cp.async.bulk + mbarrier instead of 64 cp.async (A/B variant,
  // option "tma_gather"; instantiated for the headline combination only)
  template <int MAT, bool STIFF, int TECH, bool TMA = false>
  __global__ void __launch_bounds__((TECH == 1 ? HEX8_FBAR_WARPS : HEX8_WARPS) * 32, TECH == 1 ? 1 : 2)
      k_hex8_eval_assemble(const EvalArgs A)
  {
    constexpr bool FBAR = TECH == 1, MULF = TECH == 2, LINEAR = TECH == 3;
    constexpr int WARPS = FBAR ? HEX8_FBAR_WARPS : HEX8_WARPS;
    constexpr size_t warp_bytes = FBAR ? hex8_fbar_warp_bytes : hex8_warp_bytes;
    constexpr int MP_SIZE = FBAR ? hex8fbar::MP_SIZE : hex8::MP_SIZE;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* tab_gp = reinterpret_cast<double*>(smem_raw);
    double* tab_node = tab_gp + hex8::TAB_SIZE;
    double* wgt = tab_node + hex8::TAB_SIZE;
    for (int t = threadIdx.x; t < 8 * 24; t += blockDim.x)
    {
      tab_gp[(t / 24) * hex8::TAB_STRIDE + (t % 24)] = c_hex8.dN_gp[t];
      tab_node[(t / 24) * hex8::TAB_STRIDE + (t % 24)] = c_hex8.dN_node[t];
    }
    if (threadIdx.x < 8) wgt[threadIdx.x] = c_hex8.w[threadIdx.x];
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char* wbase = smem_raw + hex8_table_bytes + warp * warp_bytes;
    hex8::WarpSmem S;
    S.mp = reinterpret_cast<double*>(wbase);
    S.g = S.mp + MP_SIZE;
    S.xu = S.g + hex8::G_SIZE;
    if (FBAR) S.h0 = S.xu + hex8::XU_SIZE;
    S.nrow = reinterpret_cast<NodeRow*>(S.xu + hex8::XU_SIZE + (FBAR ? hex8fbar::H0_SIZE : 0));
    S.pos = reinterpret_cast<int32_t*>(S.nrow + hex8::EPW * hex8::NN);
    // the warp's mbarrier (TMA variant) lives in the 16 spare bytes behind the scatter map
    uint64_t* bar = reinterpret_cast<uint64_t*>(S.pos + hex8::EPW * hex8::NN * hex8::NN);
    unsigned bar_parity = 0;
    if (TMA)
    {
      if (lane == 0) mbar_init(bar, 1);
      __syncwarp();
    }

    const int64_t n_tiles = (A.n_ele + hex8::EPW - 1) / hex8::EPW;
    const int64_t wstride = (int64_t)gridDim.x * WARPS;
    double detJ_min = 1e300, detF_min = 1e300;
    long long bad = LLONG_MAX;
    unsigned int nbad = 0;

    // prologue of the gather pipeline: tile 0 in flight, node ids of tile 1 in registers
    int64_t tile = (int64_t)blockIdx.x * WARPS + warp;
    int node_next = hex8::load_node(A, lane, tile, n_tiles);
    // Two commit groups per tile, always in the order [X,u] then [row info, scatter map]; a group
    // may be empty (no next tile), which keeps the wait_group arithmetic uniform.
    if (tile < n_tiles) hex8::issue_xu(A, S, lane, node_next);
    cp_async_commit();
    if (tile < n_tiles) hex8::issue_rowpos<TMA>(A, S, lane, node_next, tile, bar);
    cp_async_commit();
    node_next = hex8::load_node(A, lane, tile + wstride, n_tiles);

    for (; tile < n_tiles; tile += wstride)
    {
      const int64_t e0 = tile * hex8::EPW;
      cp_async_wait_but_last();  // X,u of this tile landed (its row info / map may still fly)
      __syncwarp();
      // node ids two tiles ahead: consumed at the bottom of the next iteration
      const int node_next2 = hex8::load_node(A, lane, tile + 2 * wstride, n_tiles);
      double out54[54];
      double dJ, dF;
      bool neg;
      if (FBAR)
        hex8fbar::phase1<MAT, STIFF>(A, S, tab_gp, wgt, lane, e0, out54, dJ, dF, neg);
      else
        hex8::phase1<MAT, STIFF, MULF, LINEAR>(A, S, tab_gp, tab_node, wgt, lane, e0, out54, dJ, dF, neg);
      if (e0 + (lane >> 3) < A.n_ele)
      {
        detJ_min = fmin(detJ_min, dJ);
        detF_min = fmin(detF_min, dF);
      }
      const unsigned negmask = __ballot_sync(0xffffffffu, neg);
      if (negmask)  // FOUR_C_ASSERT_ALWAYS(det > 0) of the reference: count elements, keep the first
        for (int el = 0; el < hex8::EPW; ++el)
          if ((negmask >> (8 * el)) & 0xffu)
          {
            ++nbad;
            const long long id = A.ele_id ? (long long)A.ele_id[e0 + el] : (long long)(e0 + el);
            bad = min(bad, id);
          }
      __syncwarp();  // every lane is done with the gathered X/u: refill it for the next tile
      if (tile + wstride < n_tiles) hex8::issue_xu(A, S, lane, node_next);
      cp_async_commit();
      hex8::phase1_store<STIFF>(S, lane, out54);
      __syncwarp();
      double Kacc[FBAR ? 8 : 5][9], facc[3];
      if (FBAR)
        hex8fbar::phase2<STIFF>(S, lane, Kacc, facc);
      else
        hex8::phase2<STIFF>(S, lane, Kacc, facc);
      __syncwarp();  // M/Pw/G are dead: K_e / f_e staging overwrites them
      if (FBAR)
        hex8fbar::stage_ke<STIFF>(S, lane, Kacc, facc);
      else
        hex8::stage_ke<STIFF>(S, lane, Kacc, facc);
      cp_async_wait_but_last();  // row info / scatter map of this tile landed
      if (TMA)
      {
        mbar_wait(bar, bar_parity);  // ... the map through the bulk copy's transaction count
        bar_parity ^= 1u;
      }
      __syncwarp();
      if (A.scatter_mode == 0) hex8::scatter<STIFF>(A, S, lane, e0);
      __syncwarp();  // row info / scatter map consumed: refill for the next tile
      if (tile + wstride < n_tiles) hex8::issue_rowpos<TMA>(A, S, lane, node_next, tile + wstride, bar);
      cp_async_commit();
      node_next = node_next2;
    }

    // error word: warp reduce, one atomic per warp and quantity
    for (int o = 16; o > 0; o >>= 1)
    {
      detJ_min = fmin(detJ_min, __shfl_xor_sync(0xffffffffu, detJ_min, o));
      detF_min = fmin(detF_min, __shfl_xor_sync(0xffffffffu, detF_min, o));
    }
    if (lane == 0)
    {
      atomic_min_double(&A.err->min_detJ, detJ_min);
      atomic_min_double(&A.err->min_detF, detF_min);
      if (nbad)
      {
        atomicMin(&A.err->first_bad, bad);
        atomicAdd(&A.err->n_neg_jac, (unsigned long long)nbad);
      }
    }
  }

#endif  // B200S_PART_HEX8

  // ------------------------------------------------------------------------------------------------
  // hex8 EAS evaluate + assemble kernel (hex8_eas_kernel.cuh): one warp per element, persistent grid
  // ------------------------------------------------------------------------------------------------
  constexpr int HEX8_EAS_WARPS = 4;
  constexpr int HEX8_EAS_CTAS_PER_SM = 3;  // 12 warps / SM: 170 registers, 15 KB shared memory per warp
  template <int NEAS>
  constexpr size_t hex8_eas_smem_bytes()
  {
    return hex8_table_bytes + 24 * 8 + HEX8_EAS_WARPS * (size_t)((hex8eas::Layout<NEAS>::BYTES + 15) & ~15);
  }
#if B200S_HAS(B200S_PART_EAS)
  template <int MAT, bool STIFF, int NEAS>
  __global__ void __launch_bounds__(HEX8_EAS_WARPS * 32, HEX8_EAS_CTAS_PER_SM)
      k_hex8_eas_eval_assemble(const EvalArgs A, double* hist, const double* t0tab)
  {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* tab_gp = reinterpret_cast<double*>(smem_raw);
    double* wgt = tab_gp + 2 * hex8::TAB_SIZE;
    double* gxi = wgt + 8;
    for (int t = threadIdx.x; t < 8 * 24; t += blockDim.x)
      tab_gp[(t / 24) * hex8::TAB_STRIDE + (t % 24)] = c_hex8.dN_gp[t];
    if (threadIdx.x < 8) wgt[threadIdx.x] = c_hex8.w[threadIdx.x];
    if (threadIdx.x < 24) gxi[threadIdx.x] = c_hex8.xi_gp[threadIdx.x];
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr size_t wbytes = (hex8eas::Layout<NEAS>::BYTES + 15) & ~15;
    unsigned char* wbase = smem_raw + hex8_table_bytes + 24 * 8 + warp * wbytes;
    double* sm = reinterpret_cast<double*>(wbase);
    NodeRow* nrow = reinterpret_cast<NodeRow*>(sm + hex8eas::Layout<NEAS>::DOUBLES);
    int32_t* pos = reinterpret_cast<int32_t*>(nrow + 8);
    double detJ_min = 1e300, detF_min = 1e300;
    long long bad = LLONG_MAX;
    unsigned int nbad = 0;
    constexpr int HS = hex8eas::hist_size(NEAS);
    for (int64_t e = (int64_t)blockIdx.x * HEX8_EAS_WARPS + warp; e < A.n_ele; e += (int64_t)gridDim.x * HEX8_EAS_WARPS)
    {
      bool neg = false;
      hex8eas::element<MAT, STIFF, NEAS>(A, sm, nrow, pos, tab_gp, wgt, gxi, lane, e, hist + e * HS,
          t0tab + e * hex8eas::T0_STRIDE, detJ_min, detF_min, neg);
      if (__ballot_sync(0xffffffffu, neg))
      {
        ++nbad;
        const long long id = A.ele_id ? (long long)A.ele_id[e] : (long long)e;
        bad = min(bad, id);
      }
    }
    for (int o = 16; o > 0; o >>= 1)
    {
      detJ_min = fmin(detJ_min, __shfl_xor_sync(0xffffffffu, detJ_min, o));
      detF_min = fmin(detF_min, __shfl_xor_sync(0xffffffffu, detF_min, o));
    }
    if (lane == 0)
    {
      atomic_min_double(&A.err->min_detJ, detJ_min);
      atomic_min_double(&A.err->min_detF, detF_min);
      if (nbad)
      {
        atomicMin(&A.err->first_bad, bad);
        atomicAdd(&A.err->n_neg_jac, (unsigned long long)nbad);
      }
    }
  }

  // The same element on the FP64 tensor cores (hex8_eas_mma_kernel.cuh).  The gather of the next element (node ids one
  // more element ahead in a register) is in flight as cp.async copies while the warp works on the current one.
  template <int NEAS>
  struct EasMmaCfg
  {
#ifndef B200S_EASM_WARPS_MILD
#define B200S_EASM_WARPS_MILD 6
#endif
#ifndef B200S_EASM_WARPS_FULL
#define B200S_EASM_WARPS_FULL 6
#endif
    // x 2 CTAs per SM = 12 warps for both sizes (measured: eas_mild with 14 warps at 128 registers spills, 11.2 ms against
    // 10.6 ms with 12 warps at 168; eas_full with 10 warps 21.1 ms against 17.7 ms with 12)
    static constexpr int WARPS = NEAS > 12 ? B200S_EASM_WARPS_FULL : B200S_EASM_WARPS_MILD;
    static constexpr int CTAS = 2;
    static constexpr size_t WBYTES = ((size_t)hex8eas::MmaLayout<NEAS>::BYTES + 15) & ~(size_t)15;
    static constexpr size_t SMEM = hex8_table_bytes + 72 * 8 + WARPS * WBYTES;
  };
  template <int MAT, bool STIFF, int NEAS>
  __global__ void __launch_bounds__(EasMmaCfg<NEAS>::WARPS * 32, EasMmaCfg<NEAS>::CTAS)
      k_hex8_eas_mma_eval_assemble(const EvalArgs A, double* hist, const double* t0tab)
  {
    using L = hex8eas::MmaLayout<NEAS>;
    constexpr int WARPS = EasMmaCfg<NEAS>::WARPS;
    constexpr int HS = hex8eas::hist_size(NEAS);
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* tab_gp = reinterpret_cast<double*>(smem_raw);
    double* wgt = tab_gp + 2 * hex8::TAB_SIZE;
    double* gxi = wgt + 8;
    for (int t = threadIdx.x; t < 8 * 24; t += blockDim.x)
      tab_gp[(t / 24) * hex8::TAB_STRIDE + (t % 24)] = c_hex8.dN_gp[t];
    if (threadIdx.x < 8) wgt[threadIdx.x] = c_hex8.w[threadIdx.x];
    if (threadIdx.x < 24) gxi[threadIdx.x] = c_hex8.xi_gp[threadIdx.x];
    // gxi[24 + 8 m + g]: monomial m (r, s, t, rs, rt, st) of the EAS shape functions at Gauss point g
    if (threadIdx.x < 48)
    {
      const int g = threadIdx.x & 7;
      gxi[24 + threadIdx.x] = hex8eas::monomial(threadIdx.x >> 3, c_hex8.xi_gp[3 * g], c_hex8.xi_gp[3 * g + 1], c_hex8.xi_gp[3 * g + 2]);
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double* sm = reinterpret_cast<double*>(smem_raw + hex8_table_bytes + 72 * 8 + warp * EasMmaCfg<NEAS>::WBYTES);
    double detJ_min = 1e300, detF_min = 1e300;
    long long bad = LLONG_MAX;
    unsigned int nbad = 0;
    const int64_t stride = (int64_t)gridDim.x * WARPS;
    int64_t e = (int64_t)blockIdx.x * WARPS + warp;
    // stage `buf` <- element `el` whose node ids the lanes < 8 hold in `node`
    auto issue = [&](int buf, int64_t el, int32_t node)
    {
      double* st = sm + buf * L::STAGE;
      const int32_t nk = __shfl_sync(0xffffffffu, node, lane < 24 ? lane / 3 : 0);  // node of dof `lane`
      if (lane < 24)
      {
        cp_async8(st + L::ST_XU + lane, A.X + 3 * (int64_t)nk + (lane - 3 * (lane / 3)));
        cp_async8(st + L::ST_XU + 24 + lane, A.u + 3 * (int64_t)nk + (lane - 3 * (lane / 3)));
      }
      if (lane < 8) cp_async16(reinterpret_cast<NodeRow*>(st + L::ST_NROW) + lane, A.node_row + node);
      if (lane < 16) cp_async16(reinterpret_cast<int32_t*>(st + L::ST_POS) + 4 * lane, A.pos + el * 64 + 4 * lane);
      if (lane < 20) cp_async16(st + L::ST_T0 + 2 * lane, t0tab + el * hex8eas::T0_STRIDE + 2 * lane);
      if (lane < NEAS) cp_async8(st + L::ST_ALPHA + lane, hist + el * HS + NEAS + lane);
    };
    // The node ids of an element travel one more element ahead, as a cp.async into the stage that is being worked on:
    // group(e) = { stage of element e + stride, node ids of element e + 2 stride }.
    auto fetch_conn = [&](int buf, int64_t el)
    {
      if (lane < 2) cp_async16(reinterpret_cast<int32_t*>(sm + buf * L::STAGE + L::ST_CONN) + 4 * lane, A.conn + el * 8 + 4 * lane);
    };
    if (e < A.n_ele)
    {
      const int32_t node0 = lane < 8 ? A.conn[e * 8 + lane] : 0;
      issue(0, e, node0);
      if (e + stride < A.n_ele) fetch_conn(1, e + stride);
    }
    cp_async_commit();
    // Mtilde column 16 r + (lane & 15) of this lane: Voigt row and monomial of its non-zero (-1: the w S column, -2: idle)
    int m_row[NEAS > 12 ? 2 : 1], m_mono[NEAS > 12 ? 2 : 1];
#pragma unroll
    for (int r = 0; r < (NEAS > 12 ? 2 : 1); ++r)
    {
      const int a = 16 * r + (lane & 15);
      m_row[r] = a < NEAS ? hex8eas::c_eas_row[a < NEAS ? a : 0] : (a == NEAS ? -1 : -2);
      m_mono[r] = hex8eas::c_eas_mono[a < NEAS ? a : 0];
    }
    int buf = 0;
    for (; e < A.n_ele; e += stride, buf ^= 1)
    {
      cp_async_wait_all();  // this element's stage and the node ids of the next element landed
      __syncwarp();
      if (e + stride < A.n_ele)
      {
        const int32_t node1 = lane < 8 ? reinterpret_cast<const int32_t*>(sm + (buf ^ 1) * L::STAGE + L::ST_CONN)[lane] : 0;
        issue(buf ^ 1, e + stride, node1);
        if (e + 2 * stride < A.n_ele) fetch_conn(buf, e + 2 * stride);
      }
      cp_async_commit();
      bool neg = false;
      hex8eas::element_mma<MAT, STIFF, NEAS>(A, sm, sm + buf * L::STAGE, tab_gp, wgt, gxi, m_row, m_mono, lane, e,
          hist + e * HS, detJ_min, detF_min, neg);
      if (__ballot_sync(0xffffffffu, neg))
      {
        ++nbad;
        const long long id = A.ele_id ? (long long)A.ele_id[e] : (long long)e;
        bad = min(bad, id);
      }
    }
    for (int o = 16; o > 0; o >>= 1)
    {
      detJ_min = fmin(detJ_min, __shfl_xor_sync(0xffffffffu, detJ_min, o));
      detF_min = fmin(detF_min, __shfl_xor_sync(0xffffffffu, detF_min, o));
    }
    if (lane == 0)
    {
      atomic_min_double(&A.err->min_detJ, detJ_min);
      atomic_min_double(&A.err->min_detF, detF_min);
      if (nbad)
      {
        atomicMin(&A.err->first_bad, bad);
        atomicAdd(&A.err->n_neg_jac, (unsigned long long)nbad);
      }
    }
  }

  // EAS set-up: centroid transformation of every element (reference geometry only)
  __global__ void __launch_bounds__(128) k_eas_centroid(int64_t n_ele, const int32_t* conn, const double* X, double* t0tab)
  {
    __shared__ double aug[4][72];
    const int64_t e = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (e >= n_ele) return;  // warp-uniform
    hex8eas::centroid_transformation(X, conn + e * 8, aug[warp], lane, t0tab + e * hex8eas::T0_STRIDE);
  }

  // struct_calc_recover, default step (SolidEleCalc::recover, 4C_solid_ele_calc.cpp:284-314; update_alpha,
  // 4C_solid_ele_calc_eas_helpers.hpp:166-190): alpha_inc = -invKaa (s + Kda^T du), alpha += step alpha_inc.
  // One warp per element, lane a = EAS parameter; du = "residual displacement" on the column map.
  template <int NEAS>
  __global__ void __launch_bounds__(128) k_eas_recover(int64_t n_ele, const int32_t* conn, const double* du_col,
      double step_length, double* hist)
  {
    const int64_t e = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (e >= n_ele) return;
    double* h = hist + e * hex8eas::hist_size(NEAS);
    double* alpha_inc = h;
    double* alpha = h + NEAS;
    const double* s = h + 2 * NEAS;
    const double* invKaa = h + 3 * NEAS;
    const double* Kda = invKaa + NEAS * NEAS;
    // du of the element's 24 dofs: lane r < 24
    double du = 0.0;
    if (lane < 24) du = du_col[3 * (int64_t)conn[e * 8 + lane / 3] + lane % 3];
    double tmp = 0.0;  // (s + Kda^T du)(a) in lane a
    for (int r = 0; r < 24; ++r)
    {
      const double dr = __shfl_sync(0xffffffffu, du, r);
      if (lane < NEAS) tmp = fma(Kda[r + 24 * lane], dr, tmp);
    }
    if (lane < NEAS) tmp += s[lane];
    double inc = 0.0;
    for (int c = 0; c < NEAS; ++c)
    {
      const double tc = __shfl_sync(0xffffffffu, tmp, c);
      if (lane < NEAS) inc = fma(invKaa[lane + NEAS * c], tc, inc);
    }
    if (lane < NEAS)
    {
      alpha_inc[lane] = -inc;
      alpha[lane] += step_length * (-inc);
    }
  }

#endif  // B200S_PART_EAS

  // ------------------------------------------------------------------------------------------------
  // tet10 evaluate + assemble kernel (see tet10_kernel.cuh; same gather pipeline as hex8)
  // ------------------------------------------------------------------------------------------------
  constexpr int TET10_WARPS = 4;
  constexpr size_t tet10_warp_bytes = (size_t)tet10::WARP_BYTES;
  constexpr size_t tet10_table_bytes = (size_t)(tet10::TAB_GP_SIZE + tet10::TAB_NODE_SIZE + 4) * 8;
  constexpr size_t tet10_smem_bytes = tet10_table_bytes + TET10_WARPS * tet10_warp_bytes;
  static_assert(tet10_warp_bytes % 16 == 0 && tet10_table_bytes % 16 == 0, "16-byte aligned carve-up");

  namespace tet10
  {
    __device__ __forceinline__ int load_node(const EvalArgs& A, int lane, int64_t tile, int64_t n_tiles)
    {
      const int el = lane / NN;
      if (tile >= n_tiles || el >= EPW) return -1;
      const int64_t e = tile * EPW + el;
      return e < A.n_ele ? __ldg(A.conn + e * NN + (lane - el * NN)) : -1;
    }
    __device__ __forceinline__ void issue_xu(const EvalArgs& A, const WarpSmem& S, int lane, int node)
    {
      if (node < 0)
      {
        gather_pad(S, lane);
        return;
      }
      const int el = lane / NN, q = lane - el * NN;
      double* xu = S.xu + el * XU_ESTRIDE + 3 * q;
      const double* gx = A.X + 3 * (int64_t)node;
      const double* gu = A.u + 3 * (int64_t)node;
#pragma unroll
      for (int d = 0; d < 3; ++d)
      {
        cp_async8(xu + d, gx + d);
        cp_async8(xu + NDOF + d, gu + d);
      }
    }
    __device__ __forceinline__ void issue_rowpos(
        const EvalArgs& A, const WarpSmem& S, int lane, int node, int64_t tile)
    {
      if (node >= 0)
        cp_async16(S.nrow + lane, A.node_row + node);
      else
        S.nrow[lane].rowstart = -1, S.nrow[lane].rowlen = 0, S.nrow[lane].frow = -1;
      const int64_t e0 = tile * EPW;
#pragma unroll
      for (int h = 0; h < 3; ++h)
      {
        const int chunk = lane + 32 * h;  // 16-byte chunk of the tile's 1200-byte map; 25 chunks / element
        if (chunk < POS_INTS / 4 && e0 + chunk / 25 < A.n_ele)
          cp_async16_stream(S.pos + 4 * chunk, A.pos + e0 * (NN * NN) + 4 * chunk);
      }
    }
  }  // namespace tet10

#if B200S_HAS(B200S_PART_TET10)
  // FORM 0: displacement based, 1: MULF prestress, 2: KINEM linear
  template <int MAT, bool STIFF, int FORM>
  __global__ void __launch_bounds__(TET10_WARPS * 32, 2) k_tet10_eval_assemble(const EvalArgs A)
  {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* tab_gp = reinterpret_cast<double*>(smem_raw);
    double* tab_node = tab_gp + tet10::TAB_GP_SIZE;
    double* wgt = tab_node + tet10::TAB_NODE_SIZE;
    for (int t = threadIdx.x; t < 10 * 30; t += blockDim.x)
    {
      if (t < 4 * 30) tab_gp[(t / 30) * tet10::TAB_STRIDE + (t % 30)] = c_tet10.dN_gp[t];
      tab_node[(t / 30) * tet10::TAB_STRIDE + (t % 30)] = c_tet10.dN_node[t];
    }
    if (threadIdx.x < 4) wgt[threadIdx.x] = c_tet10.w[threadIdx.x];
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char* wbase = smem_raw + tet10_table_bytes + warp * tet10_warp_bytes;
    tet10::WarpSmem S;
    S.mp = reinterpret_cast<double*>(wbase);
    S.g = S.mp + tet10::MP_SIZE;
    S.xu = S.mp + tet10::STAGE_SIZE;
    S.nrow = reinterpret_cast<NodeRow*>(S.xu + tet10::XU_SIZE);
    S.pos = reinterpret_cast<int32_t*>(S.nrow + 32);

    const int64_t n_tiles = (A.n_ele + tet10::EPW - 1) / tet10::EPW;
    const int64_t wstride = (int64_t)gridDim.x * TET10_WARPS;
    const int my_el = lane / tet10::NN;
    double detJ_min = 1e300, detF_min = 1e300;
    long long bad = LLONG_MAX;
    unsigned int nbad = 0;

    int64_t tile = (int64_t)blockIdx.x * TET10_WARPS + warp;
    int node_next = tet10::load_node(A, lane, tile, n_tiles);
    if (tile < n_tiles) tet10::issue_xu(A, S, lane, node_next);
    cp_async_commit();
    if (tile < n_tiles) tet10::issue_rowpos(A, S, lane, node_next, tile);
    cp_async_commit();
    node_next = tet10::load_node(A, lane, tile + wstride, n_tiles);

    for (; tile < n_tiles; tile += wstride)
    {
      const int64_t e0 = tile * tet10::EPW;
      cp_async_wait_but_last();
      __syncwarp();
      const int node_next2 = tet10::load_node(A, lane, tile + 2 * wstride, n_tiles);
      const bool live = my_el < tet10::EPW && e0 + my_el < A.n_ele;
      // nodal Jacobian check: lane = (element, node)
      const double dJn = tet10::node_jacobian_det(S, tab_node, lane);
      const bool neg = live && !(dJn > 0.0);
#if B200S_TET10_SPLIT_PHASE1
      double Fg[9], Sg[6], Cg[21];
      double dJ, dF;
      tet10::phase1a<MAT, STIFF, FORM == 1, FORM == 2>(A, S, tab_gp, wgt, lane, e0, Fg, Sg, Cg, dJ, dF);
#else
      double out54[54];
      double dJ, dF;
      tet10::phase1<MAT, STIFF, FORM == 1, FORM == 2>(A, S, tab_gp, wgt, lane, e0, out54, dJ, dF);
#endif
      if (live)
      {
        detJ_min = fmin(detJ_min, fmin(dJn, dJ));
        detF_min = fmin(detF_min, dF);
      }
      const unsigned negmask = __ballot_sync(0xffffffffu, neg);
      if (negmask)
        for (int el = 0; el < tet10::EPW; ++el)
          if ((negmask >> (tet10::NN * el)) & 0x3ffu)
          {
            ++nbad;
            const long long id = A.ele_id ? (long long)A.ele_id[e0 + el] : (long long)(e0 + el);
            bad = min(bad, id);
          }
      __syncwarp();
      // (order measured: storing M / Pw BEFORE the refill below -- same SASS apart from the position of the six
      // LDGSTS -- costs 14.6 -> 23.1 ms and doubles the DRAM traffic on config 4; profiles/r02_summary.md)
      if (tile + wstride < n_tiles) tet10::issue_xu(A, S, lane, node_next);
      cp_async_commit();
#if B200S_TET10_SPLIT_PHASE1
      tet10::phase1b<STIFF, FORM == 2>(S, lane, Fg, Sg, Cg);
#else
      tet10::phase1_store<STIFF>(S, lane, out54);
#endif
      __syncwarp();
      double Kacc[tet10::NBLK][9], facc[3];
      tet10::phase2<STIFF>(S, lane, Kacc, facc);
      __syncwarp();
      tet10::stage_ke<STIFF>(S, lane, Kacc, facc);
      cp_async_wait_but_last();
      __syncwarp();
      if (A.scatter_mode == 0) tet10::scatter<STIFF>(A, S, lane, e0);
      __syncwarp();
      if (tile + wstride < n_tiles) tet10::issue_rowpos(A, S, lane, node_next, tile + wstride);
      cp_async_commit();
      node_next = node_next2;
    }

    for (int o = 16; o > 0; o >>= 1)
    {
      detJ_min = fmin(detJ_min, __shfl_xor_sync(0xffffffffu, detJ_min, o));
      detF_min = fmin(detF_min, __shfl_xor_sync(0xffffffffu, detF_min, o));
    }
    if (lane == 0)
    {
      atomic_min_double(&A.err->min_detJ, detJ_min);
      atomic_min_double(&A.err->min_detF, detF_min);
      if (nbad)
      {
        atomicMin(&A.err->first_bad, bad);
        atomicAdd(&A.err->n_neg_jac, (unsigned long long)nbad);
      }
    }
  }

#endif  // B200S_PART_TET10

  // ------------------------------------------------------------------------------------------------
  // hex27 evaluate + assemble kernel (see hex27_kernel.cuh): CTA = 4 warp pairs, one element per pair
  // ------------------------------------------------------------------------------------------------
  constexpr int HEX27_PAIRS = 4;
  constexpr size_t hex27_table_bytes = (size_t)(hex27::TAB_SIZE + 1 + 28) * 8 + sizeof(hex27::IntTables);
  constexpr size_t hex27_smem_bytes = hex27_table_bytes + HEX27_PAIRS * (size_t)hex27::PAIR_BYTES;
  static_assert(hex27_table_bytes % 16 == 0, "16-byte aligned carve-up");
  static_assert(hex27_smem_bytes <= 232448, "fits the 227 KB of dynamic shared memory per CTA");

  namespace hex27
  {
    __device__ __forceinline__ void pair_barrier(int pair)
    {
      asm volatile("bar.sync %0, 64;" ::"r"(1 + pair) : "memory");
    }
    // pt = thread index inside the pair (0..63); threads 0..26 own node pt
    __device__ __forceinline__ int load_node(const EvalArgs& A, int pt, int64_t e)
    {
      return (pt < NN && e < A.n_ele) ? __ldg(A.conn + e * NN + pt) : -1;
    }
    __device__ __forceinline__ void issue_xu(const EvalArgs& A, const PairSmem& S, int pt, int node)
    {
      if (node < 0) return;
      double* xu = S.xu + 3 * pt;
      const double* gx = A.X + 3 * (int64_t)node;
      const double* gu = A.u + 3 * (int64_t)node;
#pragma unroll
      for (int d = 0; d < 3; ++d)
      {
        cp_async8(xu + d, gx + d);
        cp_async8(xu + NDOF + d, gu + d);
      }
    }
    template <bool PAD>
    __device__ __forceinline__ void issue_rowpos(
        const EvalArgs& A, const PairSmem& S, int pt, int node, int64_t e)
    {
      if (node >= 0)
      {
        if (PAD && pt >= A.nn_real)  // padding node of a hex20 block: owns no row
          S.nrow[pt] = NodeRow{-1, 0, -1};
        else
          cp_async16(S.nrow + pt, A.node_row + node);
      }
      if (e >= A.n_ele) return;
#pragma unroll
      for (int h = 0; h < 3; ++h)
      {
        const int chunk = pt + 64 * h;  // 183 chunks of 16 bytes
        if (chunk < POS_STRIDE / 4) cp_async16(S.pos + 4 * chunk, A.pos + e * POS_STRIDE + 4 * chunk);
      }
    }
  }  // namespace hex27

#if B200S_HAS(B200S_PART_HEX27)
  // FORM 0: displacement based, 1: MULF prestress, 2: KINEM linear.  PAD: a hex20 block (tables from EvalArgs::tab_override,
  // element nodes >= EvalArgs::nn_real are padding)
  template <int MAT, bool STIFF, int FORM, bool PAD = false>
  __global__ void __launch_bounds__(HEX27_PAIRS * 64, 1) k_hex27_eval_assemble(const EvalArgs A)
  {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* tab_gp = reinterpret_cast<double*>(smem_raw);
    double* wgt = tab_gp + hex27::TAB_SIZE + 1;
    hex27::IntTables* itab = reinterpret_cast<hex27::IntTables*>(wgt + 28);
    int* flags = reinterpret_cast<int*>(itab->abc + 28);  // 4 spare ints of the table = per-pair flags
    // hex27 tables, or the padded tables of a hex20 block (EvalArgs::tab_override)
    const Hex27Const& tabs = PAD ? *static_cast<const Hex27Const*>(A.tab_override) : c_hex27;
    for (int t = threadIdx.x; t < hex27::TAB_SIZE; t += blockDim.x) tab_gp[t] = tabs.dN_gp[t];
    if (threadIdx.x < 28) wgt[threadIdx.x] = tabs.w[threadIdx.x];
    if (threadIdx.x < 32)
    {
      itab->abc[threadIdx.x] = threadIdx.x < 28 ? tabs.it.abc[threadIdx.x] : 0;
      itab->node_of[threadIdx.x] = tabs.it.node_of[threadIdx.x];
    }
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int pair = warp >> 1, w = warp & 1, pt = (w << 5) | lane;
    const hex27::PairSmem S =
        hex27::carve(smem_raw + hex27_table_bytes + (size_t)pair * hex27::PAIR_BYTES);
    double* stage = S.stage + w * hex27::STAGE_SIZE;

    const int64_t stride = (int64_t)gridDim.x * HEX27_PAIRS;
    double detJ_min = 1e300, detF_min = 1e300;
    long long bad = LLONG_MAX;
    unsigned int nbad = 0;

    // Gauss point / node-check role of this lane in phase 1
    const int my_gp = w == 0 ? (lane < hex27::GP_W0 ? lane : -1)
                             : (lane < hex27::NGP - hex27::GP_W0 ? hex27::GP_W0 + lane : -1);
    const int my_chk = w == 0 ? (lane >= hex27::GP_W0 ? lane - hex27::GP_W0 : -1)
                              : (lane >= 13 && 18 + lane - 13 < hex27::NN ? 18 + lane - 13 : -1);

    int64_t e = (int64_t)blockIdx.x * HEX27_PAIRS + pair;
    int node_next = hex27::load_node(A, pt, e);
    hex27::issue_xu(A, S, pt, node_next);
    cp_async_commit();
    hex27::issue_rowpos<PAD>(A, S, pt, node_next, e);
    cp_async_commit();
    node_next = hex27::load_node(A, pt, e + stride);

    for (; e < A.n_ele; e += stride)
    {
      cp_async_wait_but_last();
      hex27::pair_barrier(pair);  // X,u of this element visible to both warps
      const int node_next2 = hex27::load_node(A, pt, e + 2 * stride);
      bool neg = false;
      if (my_gp >= 0)
      {
        double out54[54];
        double dJ, dF;
        hex27::phase1<MAT, STIFF, FORM == 1, FORM == 2>(A, S, tab_gp, wgt, my_gp, e, out54, dJ, dF);
        detJ_min = fmin(detJ_min, dJ);
        detF_min = fmin(detF_min, dF);
      }
      else if (!PAD && my_chk >= 0)  // (hex20 blocks: k_nodal_jacobian_check ahead of this kernel)
      {
        const double dJn = hex27::node_jacobian_det(S, *itab, my_chk);
        detJ_min = fmin(detJ_min, dJn);
        neg = !(dJn > 0.0);
      }
      if (__any_sync(0xffffffffu, neg) && lane == 0) atomicOr(&flags[pair], 1);
      hex27::pair_barrier(pair);  // dN/dX, M, Pw complete; X,u free for the next element
      if (pt == 0 && flags[pair])
      {
        flags[pair] = 0;
        ++nbad;
        const long long id = A.ele_id ? (long long)A.ele_id[e] : (long long)e;
        bad = min(bad, id);
      }
      hex27::issue_xu(A, S, pt, node_next);
      cp_async_commit();

      double Kacc[hex27::NBLK][9], facc[3];
      if (lane < hex27::NN) hex27::phase2<STIFF>(S, w, lane, Kacc, facc);
      cp_async_wait_but_last();
      hex27::pair_barrier(pair);  // row info / scatter map of this element visible to both warps
      if (A.scatter_mode == 0)
      {
        if (STIFF)
        {
#pragma unroll
          for (int s = 0; s < 2; ++s)
          {
            if (lane < hex27::NN) hex27::stage_blocks(stage, lane, s, Kacc);
            __syncwarp();
            hex27::scatter_blocks<PAD>(A, S, stage, w, s, lane);
            __syncwarp();
          }
        }
        if (w == 0 && lane < hex27::NN) hex27::scatter_force(A, S, lane, facc);
      }
      hex27::pair_barrier(pair);  // everything of this element consumed
      hex27::issue_rowpos<PAD>(A, S, pt, node_next, e + stride);
      cp_async_commit();
      node_next = node_next2;
    }

    for (int o = 16; o > 0; o >>= 1)
    {
      detJ_min = fmin(detJ_min, __shfl_xor_sync(0xffffffffu, detJ_min, o));
      detF_min = fmin(detF_min, __shfl_xor_sync(0xffffffffu, detF_min, o));
    }
    if (lane == 0)
    {
      atomic_min_double(&A.err->min_detJ, detJ_min);
      atomic_min_double(&A.err->min_detF, detF_min);
      if (nbad)
      {
        atomicMin(&A.err->first_bad, bad);
        atomicAdd(&A.err->n_neg_jac, (unsigned long long)nbad);
      }
    }
  }

#endif  // B200S_PART_HEX27

#if B200S_HAS(B200S_PART_CORE)
  // ------------------------------------------------------------------------------------------------
  // set-up kernels
  // ------------------------------------------------------------------------------------------------
  // ---- locality ordering of the elements (set_mesh) ----------------------------------------------
  // The assembly result does not depend on the order in which elements are visited (up to rounding
  // of the sums); the device visits them along a Morton curve of their centroids so that the CSR
  // rows a neighbourhood of elements adds to stay L2 resident until they are complete, whatever
  // numbering the caller's mesh has (structured lexicographic, or randomly renumbered tets).
  __device__ __forceinline__ void atomic_max_double(double* addr, double v)
  {
    unsigned long long* a = reinterpret_cast<unsigned long long*>(addr);
    unsigned long long old = *a;
    while (__longlong_as_double((long long)old) < v)
    {
      const unsigned long long assumed = old;
      old = atomicCAS(a, assumed, (unsigned long long)__double_as_longlong(v));
      if (old == assumed) break;
    }
  }
  __global__ void k_bbox(int64_t n_nodes, const double* X, double* box /*min xyz, max xyz*/)
  {
    double mn[3] = {1e300, 1e300, 1e300}, mx[3] = {-1e300, -1e300, -1e300};
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n_nodes;
         k += (int64_t)gridDim.x * blockDim.x)
      for (int d = 0; d < 3; ++d)
      {
        const double x = X[3 * k + d];
        mn[d] = fmin(mn[d], x), mx[d] = fmax(mx[d], x);
      }
    for (int d = 0; d < 3; ++d)
    {
      for (int o = 16; o > 0; o >>= 1)
      {
        mn[d] = fmin(mn[d], __shfl_xor_sync(0xffffffffu, mn[d], o));
        mx[d] = fmax(mx[d], __shfl_xor_sync(0xffffffffu, mx[d], o));
      }
      if ((threadIdx.x & 31) == 0)
      {
        atomic_min_double(box + d, mn[d]);
        atomic_max_double(box + 3 + d, mx[d]);
      }
    }
  }
  __device__ __forceinline__ unsigned long long spread21(unsigned long long v)
  {
    v &= 0x1fffffull;
    v = (v | (v << 32)) & 0x1f00000000ffffull;
    v = (v | (v << 16)) & 0x1f0000ff0000ffull;
    v = (v | (v << 8)) & 0x100f00f00f00f00full;
    v = (v | (v << 4)) & 0x10c30c30c30c30c3ull;
    v = (v | (v << 2)) & 0x1249249249249249ull;
    return v;
  }
  // Hilbert index of a point with 21-bit integer coordinates (Skilling's transpose form, then interleaved): a
  // space-filling curve without the long jumps of the Z-order curve -- the set of partially assembled CSR rows behind the
  // evaluation front stays smaller (option reorder = 2; A/B in profiles/r02z_summary.md)
  __device__ __forceinline__ unsigned long long hilbert_key21(unsigned int x0, unsigned int x1, unsigned int x2)
  {
    unsigned int X[3] = {x0, x1, x2};
    const unsigned int M = 1u << 20;
    for (unsigned int Q = M; Q > 1; Q >>= 1)
    {
      const unsigned int P = Q - 1;
#pragma unroll
      for (int i = 0; i < 3; ++i)
      {
        if (X[i] & Q)
          X[0] ^= P;
        else
        {
          const unsigned int t = (X[0] ^ X[i]) & P;
          X[0] ^= t, X[i] ^= t;
        }
      }
    }
    X[1] ^= X[0], X[2] ^= X[1];
    unsigned int t = 0;
    for (unsigned int Q = M; Q > 1; Q >>= 1)
      if (X[2] & Q) t ^= Q - 1;
    X[0] ^= t, X[1] ^= t, X[2] ^= t;
    return (spread21(X[0]) << 2) | (spread21(X[1]) << 1) | spread21(X[2]);
  }
  __global__ void k_morton_keys(int nn, int64_t n_ele, const int32_t* conn, const double* X,
      const double* box, unsigned long long* keys, int32_t* idx, int hilbert)
  {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_ele) return;
    double c[3] = {0, 0, 0};
    const int nc = nn < 8 ? nn : 8;  // corner / first nodes are enough for a locality key
    for (int k = 0; k < nc; ++k)
      for (int d = 0; d < 3; ++d) c[d] += X[3 * (int64_t)conn[e * nn + k] + d];
    unsigned long long key = 0;
    unsigned int q[3];
    for (int d = 0; d < 3; ++d)
    {
      const double ext = box[3 + d] - box[d];
      double t = ext > 0 ? (c[d] / nc - box[d]) / ext : 0.0;
      t = fmin(fmax(t, 0.0), 1.0);
      q[d] = (unsigned int)(t * 2097151.0);
      key |= spread21((unsigned long long)q[d]) << d;
    }
    keys[e] = hilbert ? hilbert_key21(q[2], q[1], q[0]) : key;
    idx[e] = (int32_t)e;
  }
  __global__ void k_permute_conn(int nn, int64_t n_ele, const int32_t* conn_in, const int32_t* idx,
      int32_t* conn_out)
  {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_ele * nn) return;
    conn_out[t] = conn_in[(int64_t)idx[t / nn] * nn + (t % nn)];
  }
  // ---- element segments (install_graph): see b200s_ctx::seg_begin ----------------------------------
  // row boundaries of P phases holding about nnz / P entries each
  __global__ void k_phase_rows(int P, int64_t n_rows, const int64_t* rowptr, int64_t* phase_row, int64_t* phase_off)
  {
    const int k = threadIdx.x;
    if (k > P) return;
    int64_t r = k == 0 ? 0 : n_rows;
    if (k > 0 && k < P)
    {
      const int64_t target = (int64_t)((double)rowptr[n_rows] * k / P);
      int64_t lo = 0, hi = n_rows;
      while (lo < hi)
      {
        const int64_t mid = (lo + hi) >> 1;
        if (rowptr[mid] < target)
          lo = mid + 1;
        else
          hi = mid;
      }
      r = lo - lo % 3;  // node-aligned where the three rows of a node are 3 r', 3 r' + 1, 3 r' + 2
    }
    phase_row[k] = r, phase_off[k] = rowptr[r];
  }
  // key 0: the element holds a ghost node; 1 + phase of its lowest owned row otherwise.  n_spread counts the
  // elements whose owned nodes span more than two phases (a row numbering without spatial locality).
  __global__ void k_segment_keys(int nn, int64_t n_ele, const int32_t* conn, const NodeRow* node_row, int P,
      const int64_t* phase_row, uint32_t* keys, int32_t* idx, unsigned long long* n_spread, int64_t n_owned)
  {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_ele) return;
    int pmin = INT_MAX, pmax = -1;
    bool ghost = false;
    for (int k = 0; k < nn; ++k)
    {
      const int32_t fr = node_row[conn[e * nn + k]].frow;
      if (fr < 0 || fr >= n_owned)  // not owned (export rows of the row-element mode are ghost nodes too)
      {
        ghost = true;
        continue;
      }
      int ph = 0;
      while (ph + 1 < P && phase_row[ph + 1] <= fr) ++ph;
      pmin = ph < pmin ? ph : pmin, pmax = ph > pmax ? ph : pmax;
    }
    keys[e] = (ghost || pmax < 0) ? 0u : (uint32_t)(1 + pmin);
    idx[e] = (int32_t)e;
    if (pmax - pmin > 1) atomicAdd(n_spread, 1ull);
  }
  __global__ void k_segment_bounds(int n_seg, int64_t n, const uint32_t* sorted_keys, int64_t* seg_begin)
  {
    const int s = threadIdx.x;
    if (s > n_seg) return;
    int64_t lo = 0, hi = n;
    while (lo < hi)
    {
      const int64_t mid = (lo + hi) >> 1;
      if (sorted_keys[mid] < (uint32_t)s)
        lo = mid + 1;
      else
        hi = mid;
    }
    seg_begin[s] = lo;
  }
  // out[e] = in[perm[e]] for rows of `width` items
  template <class T>
  __global__ void k_permute_rows(int64_t n, int width, const int32_t* perm, const T* in, T* out)
  {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * width) return;
    const int64_t e = t / width;
    out[t] = in[(int64_t)perm[e] * width + (t - e * width)];
  }

  // block-upper transfer: the upper part (column node >= row node) of the three CSR rows of every row node of
  // [node0, node1) -> packed, row node by row node, dof row by dof row (3 * nu contiguous doubles each)
  __global__ void k_pack_upper(int64_t node0, int64_t node1, const int64_t* csr_start, const int32_t* nb,
      const int32_t* dp, const int64_t* up_start, const double* vals, double* packed)
  {
    const int64_t r = node0 + (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31;
    if (r >= node1) return;
    const int len = 3 * nb[r], lo = 3 * dp[r], nu3 = len - lo;
    const double* src = vals + csr_start[r] + lo;
    double* dst = packed + up_start[r];
    for (int b = 0; b < 3; ++b)
      for (int q = lane; q < nu3; q += 32) dst[b * nu3 + q] = src[(int64_t)b * len + q];
  }

  __global__ void k_gather_double(int64_t n, const double* src, const int32_t* idx, double* dst)
  {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) dst[t] = src[idx[t]];
  }

  // per column node: CSR row info derived from the reference's dof maps; validates the node-blocked
  // layout the fast path of SparseMatrix::assemble relies on (4C_linalg_sparsematrix.cpp:404-441)
  __global__ void k_node_rows(int64_t n_col_nodes, const int32_t* dof_row_lid,
      const int32_t* dof_col_lid, const int64_t* rowptr, int64_t n_rows, NodeRow* node_row,
      int32_t* node_colbase, int* flag)
  {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_col_nodes) return;
    const int32_t r0 = dof_row_lid[3 * k], r1 = dof_row_lid[3 * k + 1], r2 = dof_row_lid[3 * k + 2];
    const int32_t c0 = dof_col_lid[3 * k], c1 = dof_col_lid[3 * k + 1], c2 = dof_col_lid[3 * k + 2];
    if (c0 < 0 || c1 != c0 + 1 || c2 != c0 + 2) atomicOr(flag, 1);
    node_colbase[k] = c0;
    NodeRow nr;
    nr.rowstart = -1, nr.rowlen = 0, nr.frow = -1;
    if (r0 < 0)
    {
      if (r1 >= 0 || r2 >= 0) atomicOr(flag, 2);
      node_row[k] = nr;
      return;
    }
    if (r1 != r0 + 1 || r2 != r0 + 2 || r2 >= n_rows)
    {
      atomicOr(flag, 2);
      node_row[k] = nr;
      return;
    }
    const int64_t s0 = rowptr[r0], s1 = rowptr[r1], s2 = rowptr[r2], s3 = rowptr[r2 + 1];
    if (s1 - s0 != s2 - s1 || s2 - s1 != s3 - s2) atomicOr(flag, 4);
    nr.rowstart = s0, nr.rowlen = (int32_t)(s1 - s0), nr.frow = r0;
    node_row[k] = nr;
  }

  // connectivity of a block padded from nn to nn_pad nodes per element (padding = node 0 of the element)
  __global__ void k_pad_conn(int64_t n_ele, int nn, int nn_pad, const int32_t* conn, int32_t* out)
  {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_ele * nn_pad) return;
    const int64_t e = t / nn_pad;
    const int q = (int)(t - e * nn_pad);
    out[t] = conn[e * nn + (q < nn ? q : 0)];
  }

  // ensure_positive_jacobian_determinant_at_element_nodes (4C_solid_ele_calc_lib.hpp:543-562) from the tabulated shape
  // derivatives at the nodes, one thread per element; offending elements are counted once (hex20 blocks ahead of the hex27
  // kernel, whose own check uses the Lagrange structure of hex27)
  __global__ void k_nodal_jacobian_check(int nn, int64_t n_ele, const int32_t* conn, const double* X, const double* dN_node,
      const int32_t* ele_id, ErrWord* err)
  {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_ele) return;
    const int32_t* ce = conn + e * nn;
    bool neg = false;
    double dmin = 1e300;
    for (int q = 0; q < nn; ++q)
    {
      double J[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
      const double* dN = dN_node + (size_t)q * nn * 3;
      for (int k = 0; k < nn; ++k)
      {
        const double* x = X + 3 * (int64_t)ce[k];
        for (int a = 0; a < 3; ++a)
          for (int b = 0; b < 3; ++b) J[a + 3 * b] = fma(dN[3 * k + a], x[b], J[a + 3 * b]);
      }
      const double dj = det3(J);
      dmin = fmin(dmin, dj);
      neg = neg || !(dj > 0.0);
    }
    atomic_min_double(&err->min_detJ, dmin);
    if (neg)
    {
      atomicMin(&err->first_bad, ele_id ? (long long)ele_id[e] : (long long)e);
      atomicAdd(&err->n_neg_jac, 1ull);
    }
  }

  // element -> CSR slot map: pos[e][i][j] = offset of the 3 columns of node j inside each row of
  // node i (same lower_bound + stride-3 contiguity rule as SparseMatrix::assemble :404-441)
  // nn_real < nn: the element nodes >= nn_real are padding (hex20 blocks through the hex27 kernel): entries -1
  __global__ void k_build_scatter_map(int nn, int pos_stride, int64_t n_ele, const int32_t* conn,
      const int32_t* colind, const NodeRow* node_row, const int32_t* node_colbase, int32_t* pos,
      int* flag, int nn_real)
  {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_ele * nn) return;
    const int64_t e = t / nn;
    const int i = (int)(t % nn);
    const int32_t* ce = conn + e * nn;
    const NodeRow nr = node_row[ce[i]];
    const int64_t rs = nr.rowstart;
    int32_t* out = pos + e * pos_stride + i * nn;
    if (i == 0)
      for (int t = nn * nn; t < pos_stride; ++t) pos[e * pos_stride + t] = 0;  // 16-byte padding
    if (rs < 0 || i >= nn_real)
    {
      for (int j = 0; j < nn; ++j) out[j] = -1;
      return;
    }
    const int32_t len = nr.rowlen;
    const int32_t* idx = colind + rs;
    for (int j = 0; j < nn; ++j)
    {
      if (j >= nn_real)
      {
        out[j] = -1;
        continue;
      }
      const int32_t col = node_colbase[ce[j]];
      int lo = 0, hi = len;  // lower_bound
      while (lo < hi)
      {
        const int mid = (lo + hi) >> 1;
        if (idx[mid] < col)
          lo = mid + 1;
        else
          hi = mid;
      }
      bool ok = (lo + 3 <= len);
      if (ok)
        for (int b = 0; b < 3; ++b)
          for (int c = 0; c < 3; ++c) ok = ok && (idx[(int64_t)b * len + lo + c] == col + c);
      if (!ok) atomicOr(flag, 8);
      out[j] = lo;
    }
  }

  // Discretization::set_state: owned part of the column vector (4C_fem_discretization.cpp:640-661)
  // n_owned: rows >= n_owned are export rows of ghost nodes (row-element mode), their state comes through the halo
  __global__ void k_row_to_col(int64_t n_col_nodes, const NodeRow* node_row, const double* u_row,
      double* u_col, int64_t n_owned)
  {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 3 * n_col_nodes) return;
    const int32_t fr = node_row[t / 3].frow;
    if (fr >= 0 && fr < n_owned) u_col[t] = u_row[fr + (int)(t % 3)];
  }
  __global__ void k_pack_halo(int64_t n, const int32_t* send_row_lid, const double* u_row, double* buf)
  {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 3 * n) return;
    buf[t] = u_row[send_row_lid[t / 3] + (int)(t % 3)];
  }
  __global__ void k_unpack_halo(int64_t n, const int32_t* recv_col_lid, const double* buf, double* u_col)
  {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 3 * n) return;
    u_col[3 * (int64_t)recv_col_lid[t / 3] + (t % 3)] = buf[t];
  }

  __global__ void k_count_nonfinite(int64_t n, const double* v, unsigned long long* out)
  {
    unsigned long long c = 0;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n;
         t += (int64_t)gridDim.x * blockDim.x)
      c += !isfinite(v[t]);
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, c);
  }

  // ---- graph construction (stand-in for the first assembly + FillComplete) -------------------
  __global__ void k_graph_keys(int nn, int64_t n_ele, const int32_t* conn,
      const int32_t* node_row_lid, const int32_t* node_col_lid, unsigned long long* keys)
  {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_ele * nn * nn) return;
    const int64_t e = t / (nn * nn);
    const int ij = (int)(t % (nn * nn));
    const int32_t r = node_row_lid[conn[e * nn + ij / nn]];
    const int32_t c = node_col_lid[conn[e * nn + ij % nn]];
    keys[t] = r < 0 ? ~0ull : (((unsigned long long)(uint32_t)r) << 32) | (uint32_t)c;
  }
  __global__ void k_graph_count(int64_t n_keys, const unsigned long long* keys, int* cnt)
  {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_keys || keys[t] == ~0ull) return;
    atomicAdd(&cnt[keys[t] >> 32], 1);
  }
  __global__ void k_graph_rowptr(int64_t n_row_nodes, const int64_t* nodeptr, int64_t* rowptr)
  {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r > n_row_nodes) return;
    if (r == n_row_nodes)
    {
      rowptr[3 * r] = 9 * nodeptr[r];
      return;
    }
    const int64_t c = nodeptr[r + 1] - nodeptr[r];
    for (int b = 0; b < 3; ++b) rowptr[3 * r + b] = 9 * nodeptr[r] + 3 * b * c;
  }
  __global__ void k_graph_colind(int64_t n_keys, const unsigned long long* keys,
      const int64_t* nodeptr, int32_t* colind)
  {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_keys || keys[t] == ~0ull) return;
    const int64_t r = (int64_t)(keys[t] >> 32);
    const int32_t c = (int32_t)(keys[t] & 0xffffffffu);
    const int64_t cnt = nodeptr[r + 1] - nodeptr[r], q = t - nodeptr[r];
    for (int b = 0; b < 3; ++b)
      for (int cc = 0; cc < 3; ++cc) colind[9 * nodeptr[r] + 3 * b * cnt + 3 * q + cc] = 3 * c + cc;
  }

  // ---- FP64 FMA peak ------------------------------------------------------------------------------
  __global__ void __launch_bounds__(256) k_fp64_peak(double* out, int iters, double a, double b)
  {
    double x0 = threadIdx.x * 1e-3, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5,
           x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i)
    {
#pragma unroll
      for (int r = 0; r < 8; ++r)
      {
        x0 = fma(x0, a, b), x1 = fma(x1, a, b), x2 = fma(x2, a, b), x3 = fma(x3, a, b);
        x4 = fma(x4, a, b), x5 = fma(x5, a, b), x6 = fma(x6, a, b), x7 = fma(x7, a, b);
      }
    }
    const double s = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
    if (s == 123.456) out[0] = s;
  }

  // ---- diagnostics: FP64 tensor-core MMA (DMMA m8n8k4) rate, alone and interleaved with DFMA ----
  template <int MODE>  // 1: DMMA only, 2: DMMA + DFMA interleaved
  __global__ void __launch_bounds__(256) k_dmma_peak(double* out, int iters, double a, double b)
  {
    double c0[2] = {0, 0}, c1[2] = {0, 0}, c2[2] = {0, 0}, c3[2] = {0, 0};
    double x0 = threadIdx.x * 1e-3, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3;
    const double av = a + threadIdx.x * 1e-9, bv = b + threadIdx.x * 1e-9;
    for (int i = 0; i < iters; ++i)
    {
#pragma unroll
      for (int r = 0; r < 4; ++r)
      {
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                     : "+d"(c0[0]), "+d"(c0[1]) : "d"(av), "d"(bv));
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                     : "+d"(c1[0]), "+d"(c1[1]) : "d"(av), "d"(bv));
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                     : "+d"(c2[0]), "+d"(c2[1]) : "d"(av), "d"(bv));
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                     : "+d"(c3[0]), "+d"(c3[1]) : "d"(av), "d"(bv));
        if (MODE == 2)
        {
          x0 = fma(x0, a, b), x1 = fma(x1, a, b), x2 = fma(x2, a, b), x3 = fma(x3, a, b);
          x0 = fma(x0, a, b), x1 = fma(x1, a, b), x2 = fma(x2, a, b), x3 = fma(x3, a, b);
        }
      }
    }
    const double s = c0[0] + c0[1] + c1[0] + c1[1] + c2[0] + c2[1] + c3[0] + c3[1] + x0 + x1 + x2 + x3;
    if (s == 123.456) out[0] = s;
  }

  // ---- diagnostics: red.global.add.f64 throughput for three address patterns -------------------
  // pattern 0: warp-contiguous doubles; 1: 24-byte (3 double) runs at pseudo-random row offsets inside
  // a sliding window (what the CSR scatter does); 2: fully random 8-byte
  __global__ void k_red_pattern(double* buf, int64_t n, int pattern, int per_thread)
  {
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
    for (int r = 0; r < per_thread; ++r)
    {
      const int64_t k = tid + (int64_t)r * nthreads;
      int64_t idx;
      if (pattern == 0)
        idx = k % n;
      else if (pattern == 1)
      {
        const int64_t run = k / 3;
        unsigned long long h = (unsigned long long)run * 0x9E3779B97F4A7C15ull;
        h ^= h >> 29;
        const int64_t window = 1 << 22;  // 32 MB of doubles: stays L2 resident
        const int64_t base = ((k / (3 * 4096)) * 1024) % (n - window - 4);
        idx = base + (int64_t)(h % (unsigned long long)window) / 3 * 3 + (k % 3);
      }
      else
      {
        unsigned long long h = (unsigned long long)k * 0x9E3779B97F4A7C15ull;
        h ^= h >> 31;
        idx = (int64_t)(h % (unsigned long long)n);
      }
      atomicAdd(buf + idx, 1.0);
    }
  }
  // red.global.add.f64 vs coalescing: runs of R contiguous doubles (R consecutive lanes each) at
  // pseudo-random 8-byte aligned offsets inside a sliding 32 MB (L2 resident) window; 32-bit index
  // arithmetic with compile-time R so that the ALU work stays far below the atomic rate
  template <int R>
  __global__ void k_red_runs(double* buf, unsigned n_pow2, int per_thread)
  {
    const unsigned tid = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned nthreads = gridDim.x * blockDim.x;
    const unsigned window = 1u << 22;
    for (int r = 0; r < per_thread; ++r)
    {
      const unsigned k = tid + (unsigned)r * nthreads;
      const unsigned run = k / R;
      unsigned h = run * 0x9E3779B9u;
      h ^= h >> 15;
      h *= 0x85EBCA6Bu;
      h ^= h >> 13;
      const unsigned base = ((k >> 12) * 64u) & (n_pow2 - 1);
      const unsigned idx = base + (h & (window - 1)) + (k % R);
      atomicAdd(buf + idx, 1.0);
    }
  }
#endif  // B200S_PART_CORE
}  // namespace b200s

using namespace b200s;

// --------------------------------------------------------------------------------------------------
// context
// --------------------------------------------------------------------------------------------------
struct Uid
{
  char internal[128];
};
struct NcclApi
{
  void* handle = nullptr;
  int (*GetUniqueId)(void*) = nullptr;
  int (*CommInitRank)(void**, int, /*ncclUniqueId by value (128 bytes)*/ Uid, int) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  int (*Send)(const void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*Recv)(void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};

struct SymTransfer;
struct b200s_ctx
{
  int device = 0, rank = 0, nranks = 1, n_sm = 148;
  cudaStream_t stream = nullptr;      // stream every call of this context enqueues on (an attached block: the primary's)
  cudaStream_t own_stream = nullptr;  // the stream created (and destroyed) with this context
  std::string err;
  // mesh
  int celltype = -1, nn = 0, pos_stride = 0;
  // hex20 blocks through the hex27 kernel: padded connectivity, its scatter map, padded shape tables (Hex27Const)
  int32_t *d_conn27 = nullptr, *d_pos27 = nullptr;
  void* d_tab20 = nullptr;
  int64_t n_col_nodes = 0, n_col_eles = 0;
  double* dX = nullptr;
  int32_t* dconn = nullptr;
  // SolidElementProperties of the block (4C_solid_ele_properties.hpp:55-65)
  int tech = B200S_TECH_NONE;
  int prestress = B200S_PRESTRESS_NONE;
  int kinem = B200S_KINEM_NONLINEAR_TOTLAG;
  double* d_gp_hist = nullptr;  // MULF: n_col_eles * mulf_points * MULF_HIST doubles, device element order (mulf_points = ngp,
                                // + 1 with TECH fbar: the element's global history at the centroid follows its Gauss points)
  double* d_dN_ans = nullptr;   // hex8: [8][nn][3] shape derivatives at the ANS sampling points (TECH shell_ans)
  double* d_dN_pts = nullptr;   // [ngp + 1][nn][3]: d_dN_gp followed by d_dN_cen (history set-up / update of MULF + F-bar)
  double* d_eas_hist = nullptr; // EAS: n_col_eles * hex8eas::hist_size(neas) doubles, device element order
  double* d_eas_t0 = nullptr;   // EAS: T0invT[36] + detJ0 per element (stride hex8eas::T0_STRIDE)
  double* d_hist_old = nullptr; // old_history_data_ of the block (struct_calc_update_istep / _reset_istep)
  double eas_old_step_length = 1.0;  // EASHistoryData::old_step_length (the same value in every element)
  double eas_old_step_length_snap = 1.0;  // ... of old_history_data_ (struct_calc_update_istep)
  double* d_du_row = nullptr;   // "residual displacement" state (struct_calc_recover)
  double* d_du_col = nullptr;
  // material
  int matkind = -1;
  MatParams mp{};
  double* d_per_ele = nullptr;
  HyperMat hm{};                // generic MAT_ElastHyper summand list (host copy)
  HyperMat* d_hyper = nullptr;  // ... and on the device (MatParams::hyper)
  double* d_fibers = nullptr;   // unit element fibres FIBER1, 3 per column element, device element order
  double density = 0.0;         // Mat::So3Material::density()
  double visco_in[6] = {0, 0, 0, 0, 0, 0};  // MAT_VISCONEOHOOKE input-file parameters
  double plastic_in[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};  // MAT_Struct_PlasticNlnLogNeoHooke input-file parameters
  double dt = 0.0;              // EvaluationContext::time_step_size
  bool dt_set = false;
  // cell tables in global memory for the post / mass kernels (post_kernels.cuh)
  int ngp = 0, ngp_mass = 0;
  double *d_dN_gp = nullptr, *d_w = nullptr, *d_dN_cen = nullptr, *d_xi_gp = nullptr, *d_dN_node = nullptr;
  bool general_tangent = false;  // the material's tangent has no major symmetry: generic_kernel.cuh evaluates the block
  bool general_kernel = false;   // option "general_kernel": every block through generic_kernel.cuh (A/B, tests)
  double *d_N_mass = nullptr, *d_dN_mass = nullptr, *d_w_mass = nullptr;
  uint8_t* d_ele_is_row = nullptr;  // caller numbering; nullptr = every column element is a row element
  // mass matrix / inertia force (struct_calc_nlnstiffmass & co.)
  double *d_mass = nullptr, *d_f_inertia = nullptr, *d_acc_row = nullptr, *d_acc_col = nullptr;
  // post-processing outputs
  double* d_energy = nullptr;
  // history material (history_materials.cuh): preprocessed parameters and the Gauss-point state store
  double *d_mat_ext = nullptr, *d_mat_hist = nullptr;
  // external loads (neumann_kernels.cuh)
  struct NeumannCond
  {
    bool volume = false;
    int kind = 0;
    int64_t n = 0;                 // faces / loaded elements
    int32_t* d_faces = nullptr;    // surface: [n][face_nn]
    int32_t* d_ele_slot = nullptr; // volume: caller's element id -> index, nullptr = all
    int onoff[3] = {0, 0, 0};
    double val[3] = {0, 0, 0}, fac[3] = {1, 1, 1};
    double* d_gp_fac = nullptr;
  };
  std::vector<NeumannCond> neumann;
  int face_nn = 0, face_ngp = 0;
  double *d_face_N = nullptr, *d_face_d0 = nullptr, *d_face_d1 = nullptr, *d_face_w = nullptr;
  double *d_N_gp = nullptr, *d_fext = nullptr, *d_nm_row = nullptr, *d_nm_col = nullptr;
  int* d_nm_flag = nullptr;  // set by k_neumann_surface_lin when a face pair has no graph entry
  // linear algebra on the assembled matrix (linalg_kernels.cuh)
  int64_t n_mat_cols = 0;
  bool node3_ok = false;
  int32_t* d_blkcol = nullptr;  // node-blocked layout: one column id per 3x3 block (nnz / 9)
  int32_t* d_row_diagcol = nullptr;  // matrix-local column id of every row's diagonal entry
  double *d_xmat = nullptr, *d_la_x = nullptr, *d_la_y = nullptr, *d_toggle = nullptr;
  double *d_cg[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};  // r z p Ap dinv(3x)
  double *d_cg_s = nullptr, *d_cg_hist = nullptr;
  int cg_hist_len = 0;
  cudaEvent_t ev_la[2] = {};
  double last_spmv_ms = 0.0;
  // graph
  int64_t n_rows = 0, nnz = 0;
  int64_t* d_rowptr = nullptr;  // only when built/owned here
  int32_t* d_colind = nullptr;
  int32_t* d_pos = nullptr;
  NodeRow* d_node_row = nullptr;
  int32_t* d_node_colbase = nullptr;
  int32_t* d_ele_id = nullptr;  // device element -> caller's column-element index (locality order)
  double *d_vals = nullptr, *d_f = nullptr, *d_u_row = nullptr, *d_u_col = nullptr;
  // stiff.zero() of the NEXT Newton iteration overlapped with this iteration's evaluate: two value
  // buffers, the idle one is cleared on a side stream while the element kernel fills the other
  double* d_vals_buf[2] = {nullptr, nullptr};
  static constexpr size_t ZERO_BLOCK = 64u << 20;
  void* d_zero = nullptr;         // option double_buffer = 2: source of the copy-engine zeroing
  int vals_cur = 0;               // d_vals == d_vals_buf[vals_cur]: result of the last evaluate
  bool vals_zeroed[2] = {false, false};
  cudaStream_t zero_stream = nullptr;
  cudaEvent_t ev_zero_done[2] = {}, ev_eval_start = nullptr;
  // element segments in device element order (chosen at install_graph, see order_elements_for_graph):
  //   segment 0            elements holding a ghost node: the only ones that read halo-imported displacements
  //   segments 1..n_phases the others by ROW PHASE = contiguous CSR row range (about equal nnz) of their lowest
  //                        owned node: once the segments 0..k+1 are evaluated, the rows of phase k are complete
  //                        and can travel to the host while the later segments are still being evaluated
  int n_phases = 1;
  bool segmented = false;              // false: one segment holding everything (option reorder = 0, tiny meshes)
  std::vector<int64_t> seg_begin;      // n_phases + 2 element offsets
  std::vector<int64_t> phase_row;      // n_phases + 1 row boundaries
  std::vector<int64_t> phase_off;      // rowptr[phase_row[k]]
  cudaStream_t copy_stream = nullptr, halo_stream = nullptr;
  static constexpr int MAX_PHASES = 64;
  cudaEvent_t ev_phase[MAX_PHASES] = {}, ev_fork = nullptr, ev_halo_done = nullptr, ev_copy_done = nullptr;
  // block-upper transfer of a symmetric K with a threaded host-side mirror (option symmetric_transfer)
  struct SymTransfer* sym = nullptr;
  int symmetric_transfer = 0, host_threads = 0, nt_stores = 1;
  int64_t last_d2h_bytes = 0;  // bytes of K values that crossed PCIe in the last b200s_evaluate
  // element block attached to another context's system matrix (b200s_attach_block): graph, values, f, the
  // displacement vectors, stream and communicator are the primary's (not owned here)
  b200s_ctx* primary = nullptr;
  // halo
  int n_nbr = 0;
  std::vector<int32_t> peer;
  std::vector<int64_t> send_off, recv_off;
  int32_t *d_send_lid = nullptr, *d_recv_lid = nullptr;
  double *d_sendbuf = nullptr, *d_recvbuf = nullptr;
  NcclApi nccl;
  void* comm = nullptr;
  // errors / timers
  ErrWord* d_err = nullptr;
  ErrWord* h_err = nullptr;       // pinned; D2H target only (read at b200s_synchronize)
  ErrWord* h_err_init = nullptr;  // pinned constant: the reset word every evaluate copies to d_err
  cudaEvent_t ev[7] = {};
  static constexpr int RING = 64;       // element-kernel event pairs not yet harvested
  cudaEvent_t kev[2 * RING] = {};
  bool kev_first[RING] = {};  // pair belongs to the first element-kernel launch of an evaluate
  int kev_pending = 0;
  double ktime_sum_ms = 0.0;
  int64_t ktime_count = 0;
  double timers[6] = {0, 0, 0, 0, 0, 0};
  int64_t launches = 0;
  int last_action = 0;
  bool timed_h2d = false, timed_d2h = false;
  // options
  int check_finite = 0, ctas_per_sm = 2, scatter_mode = 0, reorder = 1, double_buffer = 0;
  int segments = 1;          // 0: keep the plain locality order of set_mesh (no ghost / row-phase segments)
  int tma_gather = 0;        // hex8 headline kernel: scatter map staged by a TMA bulk copy (A/B variant)
  int eas_mma = 1;           // hex8 EAS: element matrices on the FP64 tensor cores (0: warp-per-element FMA kernel)
  int pipeline_phases = 12;  // row phases of the host-destination pipeline (1 = off)
  int overlap_halo = 1;      // ghost import on a side stream while the elements without ghost nodes are evaluated
  // row-element mode (FE_MATRIX: 4C_linalg_sparsematrix.cpp:598-630,705-710): the graph carries export rows behind
  // the rank's own ones; options "owned_rows" / "owned_nnz" (before set_graph) give the split, b200s_set_row_export
  // the peers, slices and add positions
  int64_t n_rows_owned = 0, nnz_owned = 0;  // 0: every row of the graph is the rank's own
  std::vector<int> exp_peer;
  std::vector<int64_t> exp_send_val, exp_send_f, exp_recv_off, exp_recv_foff;  // per peer {start, count} pairs / offsets
  int64_t* d_exp_pos = nullptr;   // add positions in the owned values, concatenated over the peers
  int32_t* d_exp_fpos = nullptr;  // add positions in the owned residual
  double *d_exp_buf = nullptr, *d_exp_fbuf = nullptr;  // receive buffers
  // what the last evaluate did (bench / tests): launches of the element kernel, overlapped D2H chunks
  int last_kernel_launches = 0, last_pipeline_chunks = 0, last_halo_overlapped = 0;
};

namespace
{
  // (static: nvcc derives the anonymous-namespace name from the source file, which all parts share)
  static int fail(b200s_ctx* c, int code, const char* fmt, ...)
  {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (c) c->err = buf;
    return code;
  }
#define CK(call)                                                                            \
  do                                                                                        \
  {                                                                                         \
    cudaError_t e_ = (call);                                                                \
    if (e_ != cudaSuccess)                                                                  \
      return fail(ctx, B200S_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), \
          __FILE__, __LINE__);                                                              \
  } while (0)

  template <class T>
  int to_device(b200s_ctx* ctx, T** dst, const T* src, int64_t n)
  {
    if (*dst) cudaFree(*dst);
    *dst = nullptr;
    CK(cudaMalloc((void**)dst, sizeof(T) * (size_t)(n > 0 ? n : 1)));
    if (n > 0 && src)
      CK(cudaMemcpyAsync(*dst, src, sizeof(T) * (size_t)n, cudaMemcpyDefault, ctx->stream));
    return 0;
  }
  template <class T>
  void dfree(T*& p)
  {
    if (p) cudaFree(p);
    p = nullptr;
  }
  // blocks for n work items; at least one, so that a processor without elements / row nodes (ordinary in the reference's
  // column-element loop) launches a block that finds nothing to do instead of an invalid empty grid
  inline unsigned nblk(int64_t n, int bs = 256) { return n > 0 ? (unsigned)((n + bs - 1) / bs) : 1u; }

  template <class K>
  int set_smem_attr(K kernel, size_t bytes)
  {
    return (int)cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  }
}  // namespace

// Entry points of the kernel translation units (each defined in exactly one part, see the top of the file)
namespace b200s_parts
{
  int launch_hex8(b200s_ctx* ctx, const EvalArgs& A, bool stiff);      // TECH none / fbar, MULF, KINEM linear
  int launch_hex8_eas(b200s_ctx* ctx, const EvalArgs& A, bool stiff);  // TECH eas_mild / eas_full
  int launch_tet10(b200s_ctx* ctx, const EvalArgs& A, bool stiff);
  int launch_hex27(b200s_ctx* ctx, const EvalArgs& A, bool stiff);
  int set_tables_hex8(b200s_ctx* ctx, const Hex8Const& h);      // constant tables of the part's kernels
  int set_tables_hex8_eas(b200s_ctx* ctx, const Hex8Const& h);
  int set_tables_tet10(b200s_ctx* ctx, const Tet10Const& h);
  int set_tables_hex27(b200s_ctx* ctx, const Hex27Const& h);
  int eas_centroid(b200s_ctx* ctx);                              // set-up: T0^-T, det J0 of every element
  int eas_recover(b200s_ctx* ctx, int neas, double step_length); // struct_calc_recover
}  // namespace b200s_parts

// material switch of a part's launch function template
#define B200S_DISPATCH_MATERIAL(fn)                                                                     \
  switch (ctx->matkind)                                                                                 \
  {                                                                                                     \
    case B200S_MAT_COUP_NEOHOOKE:                                                                       \
      return stiff ? fn<MAT_COUP_NEOHOOKE, true>(ctx, A) : fn<MAT_COUP_NEOHOOKE, false>(ctx, A);        \
    case B200S_MAT_ISO_VOL:                                                                             \
      return stiff ? fn<MAT_ISO_VOL, true>(ctx, A) : fn<MAT_ISO_VOL, false>(ctx, A);                    \
    case B200S_MAT_SVK:                                                                                 \
      return stiff ? fn<MAT_SVK, true>(ctx, A) : fn<MAT_SVK, false>(ctx, A);                            \
    case B200S_MAT_ELASTHYPER:                                                                          \
      return stiff ? fn<MAT_HYPER, true>(ctx, A) : fn<MAT_HYPER, false>(ctx, A);                        \
  }                                                                                                     \
  return fail(ctx, B200S_ERR_INVALID, "material not set")

#if B200S_HAS(B200S_PART_CORE)
// ---- block-upper transfer of a symmetric K (option symmetric_transfer) ------------------------------------------
// For formulations whose K_e is symmetric (everything on the device path except F-bar) only the blocks with column
// node >= row node cross PCIe (ghost columns follow the owned ones in the column map, so they are all included); a
// small pool of host threads writes the full CSR rows into the caller's array: the upper part of a row is copied,
// a lower block (i,j) is the transpose of the upper block (j,i) of an earlier row.  Sums of the same element
// contributions in another order: equal to rounding, well inside the 1e-12 parity bound.
struct HostPool
{
  std::vector<std::thread> th;
  std::mutex m;
  std::condition_variable cv, cv_done;
  std::deque<std::function<void()>> q;
  int busy = 0;
  bool stop = false;
  void start(int n)
  {
    for (int i = 0; i < n; ++i)
      th.emplace_back([this]
          {
            for (;;)
            {
              std::function<void()> f;
              {
                std::unique_lock<std::mutex> lk(m);
                cv.wait(lk, [this] { return stop || !q.empty(); });
                if (stop && q.empty()) return;
                f = std::move(q.front());
                q.pop_front();
                ++busy;
              }
              f();
              {
                std::lock_guard<std::mutex> lk(m);
                --busy;
              }
              cv_done.notify_all();
            }
          });
  }
  void submit(std::function<void()> f)
  {
    {
      std::lock_guard<std::mutex> lk(m);
      q.push_back(std::move(f));
    }
    cv.notify_one();
  }
  void wait()
  {
    std::unique_lock<std::mutex> lk(m);
    cv_done.wait(lk, [this] { return q.empty() && busy == 0; });
  }
  ~HostPool()
  {
    {
      std::lock_guard<std::mutex> lk(m);
      stop = true;
    }
    cv.notify_all();
    for (auto& t : th) t.join();
  }
};

struct SymTransfer
{
  bool usable = false;
  std::string why;               // reason when not usable
  int64_t n_row_nodes = 0, n_packed = 0;  // packed doubles
  // per row node (row lid of dof 0 = 3 r): blocks per row, index of the diagonal block, first CSR entry, first packed
  // double, first entry of the lower-block tables
  std::vector<int32_t> nb, dp;
  std::vector<int64_t> csr_start, up_start, low_start;
  // per lower block: packed offset of the transposed source block (dof row 0) and the source's row stride 3 nu_j
  std::vector<int64_t> moff;
  std::vector<int32_t> mstride;
  std::vector<int64_t> phase_node;  // row-node boundaries of the row phases
  double *d_packed = nullptr, *h_packed = nullptr;
  int64_t *d_csr_start = nullptr, *d_up_start = nullptr;
  int32_t *d_nb = nullptr, *d_dp = nullptr;
  cudaEvent_t ev_arrive[b200s_ctx::MAX_PHASES] = {};
  HostPool pool;
  int n_threads = 0;
  ~SymTransfer()
  {
    cudaFree(d_packed), cudaFree(d_csr_start), cudaFree(d_up_start), cudaFree(d_nb), cudaFree(d_dp);
    if (h_packed) cudaFreeHost(h_packed);
    for (auto& e : ev_arrive)
      if (e) cudaEventDestroy(e);
  }
};

namespace
{
  inline void store_stream(double* p, double v, bool nt)
  {
#if defined(__x86_64__)
    if (nt)
    {
      long long bits;
      memcpy(&bits, &v, 8);
      _mm_stream_si64(reinterpret_cast<long long*>(p), bits);  // no read-for-ownership of the destination line
      return;
    }
#endif
    (void)nt;
    *p = v;
  }

  // full CSR rows of the row nodes [r0, r1) from the packed upper blocks
  void sym_unpack_rows(const SymTransfer& T, int64_t r0, int64_t r1, double* K, bool nt)
  {
    const double* P = T.h_packed;
    for (int64_t r = r0; r < r1; ++r)
    {
      const int nb = T.nb[r], dp = T.dp[r], nu3 = 3 * (nb - dp), len = 3 * nb;
      const int64_t* mo = T.moff.data() + T.low_start[r];
      const int32_t* ms = T.mstride.data() + T.low_start[r];
      for (int b = 0; b < 3; ++b)
      {
        double* dst = K + T.csr_start[r] + (int64_t)b * len;
        // K_ij(b,c) = K_ji(c,b): row c of the source block row, column 3 pos + b
        for (int k = 0; k < dp; ++k)
        {
          const double* src = P + mo[k] + b;
          const int64_t st = ms[k];
          store_stream(dst + 3 * k, src[0], nt);
          store_stream(dst + 3 * k + 1, src[st], nt);
          store_stream(dst + 3 * k + 2, src[2 * st], nt);
        }
        const double* up = P + T.up_start[r] + (int64_t)b * nu3;
        double* du = dst + 3 * dp;
        for (int q = 0; q < nu3; ++q) store_stream(du + q, up[q], nt);
      }
    }
#if defined(__x86_64__)
    if (nt) _mm_sfence();
#endif
  }

  // tables of the block-upper transfer for the graph installed in ctx (once per graph)
  int sym_prepare(b200s_ctx* ctx)
  {
    delete ctx->sym;
    SymTransfer* T = ctx->sym = new SymTransfer;
    auto no = [&](const char* why)
    {
      T->usable = false, T->why = why;
      return 0;
    };
    if (!ctx->node3_ok || !ctx->d_blkcol) return no("the graph is not node-blocked");
    if (!ctx->segmented) return no("elements are not grouped by row phase (option reorder = 0)");
    if (ctx->n_rows % 3) return no("rows are not node-blocked");
    const int64_t ncn = ctx->n_col_nodes, nrn = ctx->n_rows / 3;
    std::vector<NodeRow> nr((size_t)ncn);
    std::vector<int32_t> cb((size_t)ncn), blk((size_t)(ctx->nnz / 9 + 1));
    CK(cudaMemcpy(nr.data(), ctx->d_node_row, sizeof(NodeRow) * (size_t)ncn, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(cb.data(), ctx->d_node_colbase, 4 * (size_t)ncn, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(blk.data(), ctx->d_blkcol, 4 * (size_t)(ctx->nnz / 9), cudaMemcpyDeviceToHost));
    T->n_row_nodes = nrn;
    T->nb.assign((size_t)nrn, -1), T->dp.assign((size_t)nrn, 0);
    T->csr_start.assign((size_t)nrn, 0), T->up_start.assign((size_t)nrn + 1, 0), T->low_start.assign((size_t)nrn + 1, 0);
    for (int64_t k = 0; k < ncn; ++k)
    {
      if (nr[(size_t)k].frow < 0) continue;
      const int32_t fr = nr[(size_t)k].frow;
      // owned dofs must carry their row lid as matrix column id (Epetra's column map lists the owned ids first, in
      // row-map order) and the rows of a node must be 3 r, 3 r + 1, 3 r + 2
      if (fr % 3 || cb[(size_t)k] != fr) return no("column ids of owned dofs differ from their row ids");
      const int64_t r = fr / 3;
      T->nb[(size_t)r] = nr[(size_t)k].rowlen / 3, T->csr_start[(size_t)r] = nr[(size_t)k].rowstart;
    }
    for (int64_t r = 0; r < nrn; ++r)
    {
      if (T->nb[(size_t)r] < 0) return no("a row node without CSR rows");
      const int32_t* bl = blk.data() + T->csr_start[(size_t)r] / 9;
      const int nb = T->nb[(size_t)r];
      const int dp = (int)(std::lower_bound(bl, bl + nb, (int32_t)(3 * r)) - bl);
      if (dp >= nb || bl[dp] != 3 * r) return no("a row without diagonal block");
      T->dp[(size_t)r] = dp;
      T->up_start[(size_t)r + 1] = T->up_start[(size_t)r] + 9 * (int64_t)(nb - dp);
      T->low_start[(size_t)r + 1] = T->low_start[(size_t)r] + dp;
    }
    T->n_packed = T->up_start[(size_t)nrn];
    T->moff.resize((size_t)T->low_start[(size_t)nrn]), T->mstride.resize((size_t)T->low_start[(size_t)nrn]);
    bool pattern_ok = true;
    const int nt = ctx->host_threads > 0 ? ctx->host_threads : (int)std::max(1u, std::min(32u, std::thread::hardware_concurrency()));
    {
      std::vector<std::thread> th;
      std::mutex mm;
      for (int t = 0; t < nt; ++t)
        th.emplace_back([&, t]
            {
              bool ok = true;
              for (int64_t r = nrn * t / nt; r < nrn * (t + 1) / nt; ++r)
              {
                const int32_t* bl = blk.data() + T->csr_start[(size_t)r] / 9;
                for (int k = 0; k < T->dp[(size_t)r]; ++k)
                {
                  const int64_t j = bl[k] / 3;  // owned: column id = row id
                  const int32_t* bj = blk.data() + T->csr_start[(size_t)j] / 9;
                  const int nbj = T->nb[(size_t)j], dpj = T->dp[(size_t)j];
                  const int pos = (int)(std::lower_bound(bj + dpj, bj + nbj, (int32_t)(3 * r)) - (bj + dpj));
                  if (pos >= nbj - dpj || bj[dpj + pos] != 3 * r)
                  {
                    ok = false;  // structurally unsymmetric pattern
                    continue;
                  }
                  T->moff[(size_t)(T->low_start[(size_t)r] + k)] = T->up_start[(size_t)j] + 3 * pos;
                  T->mstride[(size_t)(T->low_start[(size_t)r] + k)] = 3 * (nbj - dpj);
                }
              }
              if (!ok)
              {
                std::lock_guard<std::mutex> lk(mm);
                pattern_ok = false;
              }
            });
      for (auto& t : th) t.join();
    }
    if (!pattern_ok) return no("the sparsity pattern is not structurally symmetric");
    T->phase_node.resize(ctx->phase_row.size());
    for (size_t k = 0; k < ctx->phase_row.size(); ++k)
    {
      if (ctx->phase_row[k] % 3) return no("row phases are not node-aligned");
      T->phase_node[k] = ctx->phase_row[k] / 3;
    }
    if (to_device(ctx, &T->d_csr_start, T->csr_start.data(), nrn) || to_device(ctx, &T->d_up_start, T->up_start.data(), nrn) ||
        to_device(ctx, &T->d_nb, T->nb.data(), nrn) || to_device(ctx, &T->d_dp, T->dp.data(), nrn))
      return B200S_ERR_CUDA;
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaMalloc((void**)&T->d_packed, 8 * (size_t)std::max<int64_t>(T->n_packed, 1)));
    CK(cudaMallocHost((void**)&T->h_packed, 8 * (size_t)std::max<int64_t>(T->n_packed, 1)));
    for (auto& e : T->ev_arrive) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    T->n_threads = nt;
    T->pool.start(nt);
    T->usable = true;
    return 0;
  }

  // element-kernel durations: one CUDA event pair per evaluate on the context stream, harvested at
  // the next synchronize (or when the ring is full) so that a timed loop needs no per-step sync
  void harvest_kernel_times(b200s_ctx* ctx)
  {
    for (int i = 0; i < ctx->kev_pending; ++i)
    {
      float ms = 0;
      cudaEventSynchronize(ctx->kev[2 * i + 1]);
      if (cudaEventElapsedTime(&ms, ctx->kev[2 * i], ctx->kev[2 * i + 1]) == cudaSuccess)
      {
        // an evaluate may consist of several launches (segments): their times add up, the count is per evaluate
        ctx->ktime_sum_ms += ms, ctx->ktime_count += ctx->kev_first[i] ? 1 : 0;
        ctx->timers[2] = ctx->kev_first[i] ? ms : ctx->timers[2] + ms;
      }
    }
    ctx->kev_pending = 0;
  }

  // an attached element block borrows these from its primary: forget them before the context frees its own
  void detach_shared(b200s_ctx* ctx)
  {
    ctx->d_rowptr = nullptr, ctx->d_colind = nullptr, ctx->d_node_row = nullptr, ctx->d_node_colbase = nullptr;
    ctx->d_row_diagcol = nullptr, ctx->d_blkcol = nullptr;
    ctx->d_vals_buf[0] = ctx->d_vals_buf[1] = nullptr, ctx->d_vals = nullptr;
    ctx->d_f = nullptr, ctx->d_u_row = nullptr, ctx->d_u_col = nullptr;
    ctx->comm = nullptr, ctx->n_nbr = 0;
    ctx->stream = ctx->own_stream;
    ctx->primary = nullptr;
    ctx->n_rows = 0, ctx->nnz = 0;
  }

  // per-element arrays follow the device element order: out[e] = in[perm[e]]
  template <class T>
  static int permute_array(b200s_ctx* ctx, T*& arr, int64_t n, int width, const int32_t* d_perm)
  {
    if (!arr || n <= 0 || width <= 0) return 0;
    T* out = nullptr;
    CK(cudaMalloc((void**)&out, sizeof(T) * (size_t)n * width));
    k_permute_rows<T><<<nblk(n * width), 256, 0, ctx->stream>>>(n, width, d_perm, arr, out);
    ctx->launches += 1;
    CK(cudaStreamSynchronize(ctx->stream));
    cudaFree(arr);
    arr = out;
    return 0;
  }

  inline bool material_needs_fibers(const b200s_ctx* ctx)
  {
    if (ctx->matkind != B200S_MAT_ELASTHYPER) return false;
    for (int k = 0; k < ctx->hm.n; ++k)
      if (summand_needs_fiber(ctx->hm.s[k].kind)) return true;
    return false;
  }
  inline int eas_neas(int tech)
  {
    return tech == B200S_TECH_EAS_MILD ? 9 : (tech == B200S_TECH_EAS_FULL ? 21 : ((tech == B200S_TECH_SHELL_EAS || tech == B200S_TECH_SHELL_EAS_ANS) ? 7 : 0));
  }
  // MulfHistoryData per element: one per Gauss point, with TECH fbar one more for the centroid (GlobalHistory of
  // MulfFBarFormulation, 4C_solid_ele_calc_mulf_fbar.hpp:112-116)
  inline int mulf_points(const b200s_ctx* ctx) { return ctx->ngp + (ctx->tech == B200S_TECH_FBAR ? 1 : 0); }
  // what the caller sees of the matrix / vectors: the rank's own rows (row-element mode keeps export rows behind them)
  inline int64_t io_rows(const b200s_ctx* ctx) { return ctx->n_rows_owned > 0 ? ctx->n_rows_owned : ctx->n_rows; }
  inline int64_t io_nnz(const b200s_ctx* ctx) { return ctx->n_rows_owned > 0 ? ctx->nnz_owned : ctx->nnz; }
  // fresh EASHistoryData of every element: all zero (4C_solid_ele_calc_eas_helpers.hpp:98-112)
  int eas_allocate(b200s_ctx* ctx)
  {
    ctx->eas_old_step_length = ctx->eas_old_step_length_snap = 1.0;
    const int64_t n = ctx->n_col_eles * hex8eas::hist_size(eas_neas(ctx->tech));
    if (to_device(ctx, &ctx->d_eas_hist, (const double*)nullptr, n)) return B200S_ERR_CUDA;
    CK(cudaMemsetAsync(ctx->d_eas_hist, 0, sizeof(double) * (size_t)(n > 0 ? n : 1), ctx->stream));
    if (to_device(ctx, &ctx->d_eas_t0, (const double*)nullptr, ctx->n_col_eles * hex8eas::T0_STRIDE)) return B200S_ERR_CUDA;
    if (ctx->n_col_eles > 0 && b200s_parts::eas_centroid(ctx)) return B200S_ERR_CUDA;
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaGetLastError());
    return 0;
  }

  // history doubles per Gauss point of the block's material
  inline int mat_hist_per_gp(const b200s_ctx* ctx)
  {
    return ctx->matkind == B200S_MAT_DRUCKER_PRAGER
               ? DP_HIST
               : (ctx->matkind == B200S_MAT_VISCONEOHOOKE ? VISCO_HIST : (ctx->matkind == B200S_MAT_PLASTIC_NLNLOG ? PLASTIC_HIST : 0));
  }
  // fresh history of a history material: what Mat::PlasticDruckerPrager::setup(numgp) does per element
  // (src/mat/4C_mat_plasticdruckerprager.cpp:109-117): zero plastic strains at every Gauss point
  int material_history_allocate(b200s_ctx* ctx)
  {
    dfree(ctx->d_mat_hist);
    if (mat_hist_per_gp(ctx) == 0 || ctx->celltype < 0) return 0;
    const int64_t n = ctx->n_col_eles * ctx->ngp * mat_hist_per_gp(ctx);
    if (to_device(ctx, &ctx->d_mat_hist, (const double*)nullptr, n)) return B200S_ERR_CUDA;
    CK(cudaMemsetAsync(ctx->d_mat_hist, 0, sizeof(double) * (size_t)(n > 0 ? n : 1), ctx->stream));
    if (ctx->matkind == B200S_MAT_PLASTIC_NLNLOG && n > 0)
    {
      // PlasticNlnLogNeoHooke::setup (:337-349): inverse plastic right Cauchy-Green tensors = identity
      const int64_t np = ctx->n_col_eles * ctx->ngp;
      k_plastic_history_init<<<nblk(np, 128), 128, 0, ctx->stream>>>(np, ctx->d_mat_hist);
      ctx->launches += 1;
    }
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
  }

  // fresh MulfHistoryData of every Gauss point (inverse_jacobian of the reference configuration,
  // deformation_gradient = I): what the reference sets up lazily at the first evaluate
  // (4C_solid_ele_calc_mulf.hpp:56-60)
  int mulf_allocate(b200s_ctx* ctx)
  {
    const int64_t n = ctx->n_col_eles * mulf_points(ctx);
    if (to_device(ctx, &ctx->d_gp_hist, (const double*)nullptr, n * MULF_HIST)) return B200S_ERR_CUDA;
    if (n > 0)
    {
      k_mulf_init<<<nblk(n, 128), 128, 0, ctx->stream>>>(ctx->n_col_eles, ctx->nn, mulf_points(ctx), ctx->dconn, ctx->dX,
          ctx->d_dN_pts, ctx->d_gp_hist);
      ctx->launches += 1;
    }
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaGetLastError());
    return 0;
  }

}  // namespace
#endif  // B200S_PART_CORE

#if B200S_HAS(B200S_PART_EAS)
namespace
{
  template <int MAT, bool STIFF, int NEAS>
  static int launch_hex8_eas(b200s_ctx* ctx, const EvalArgs& A)
  {
    if (!ctx->d_eas_hist) return fail(ctx, B200S_ERR_INVALID, "EAS history missing");
    if constexpr (NEAS != 7)  // eastype_sh8_7 (TECH shell_eas): FMA kernel only
    if (ctx->eas_mma || MAT == MAT_PLASTIC)  // MAT_PLASTIC: tensor-core kernel only (one instantiation less)
    {
      auto kern = k_hex8_eas_mma_eval_assemble<MAT, STIFF, NEAS>;
      using Cfg = EasMmaCfg<NEAS>;
      if (set_smem_attr(kern, Cfg::SMEM)) return fail(ctx, B200S_ERR_CUDA, "cudaFuncSetAttribute(smem=%zu) failed", Cfg::SMEM);
      int64_t grid = (A.n_ele + Cfg::WARPS - 1) / Cfg::WARPS;
      const int64_t cap = (int64_t)ctx->n_sm * Cfg::CTAS;
      if (grid > cap) grid = cap;
      if (grid < 1) grid = 1;
      kern<<<(unsigned)grid, Cfg::WARPS * 32, Cfg::SMEM, ctx->stream>>>(A,
          ctx->d_eas_hist + A.e_begin * hex8eas::hist_size(NEAS), ctx->d_eas_t0 + A.e_begin * hex8eas::T0_STRIDE);
      ctx->launches += 1;
      return 0;
    }
    if constexpr (MAT == MAT_PLASTIC) return fail(ctx, B200S_ERR_UNSUPPORTED, "MAT_Struct_PlasticNlnLogNeoHooke with TECH shell_eas has no device kernel");
    auto kern = k_hex8_eas_eval_assemble<MAT == MAT_PLASTIC ? MAT_VISCO : MAT, STIFF, NEAS>;
    constexpr size_t bytes = hex8_eas_smem_bytes<NEAS>();
    if (set_smem_attr(kern, bytes)) return fail(ctx, B200S_ERR_CUDA, "cudaFuncSetAttribute(smem=%zu) failed", bytes);
    int64_t grid = (A.n_ele + HEX8_EAS_WARPS - 1) / HEX8_EAS_WARPS;
    const int64_t cap = (int64_t)ctx->n_sm * HEX8_EAS_CTAS_PER_SM;
    if (grid > cap) grid = cap;
    if (grid < 1) grid = 1;
    kern<<<(unsigned)grid, HEX8_EAS_WARPS * 32, bytes, ctx->stream>>>(A,
        ctx->d_eas_hist + A.e_begin * hex8eas::hist_size(NEAS), ctx->d_eas_t0 + A.e_begin * hex8eas::T0_STRIDE);
    ctx->launches += 1;
    return 0;
  }

  template <int MAT, bool STIFF>
  static int launch_hex8_eas_any(b200s_ctx* ctx, const EvalArgs& A)
  {
    if (ctx->tech == B200S_TECH_SHELL_EAS) return launch_hex8_eas<MAT, STIFF, 7>(ctx, A);
    return ctx->tech == B200S_TECH_EAS_MILD ? launch_hex8_eas<MAT, STIFF, 9>(ctx, A)
                                            : launch_hex8_eas<MAT, STIFF, 21>(ctx, A);
  }
}  // namespace
namespace b200s_parts
{
  int launch_hex8_eas(b200s_ctx* ctx, const EvalArgs& A, bool stiff)
  {
    if (ctx->matkind == B200S_MAT_VISCONEOHOOKE)
      return stiff ? launch_hex8_eas_any<MAT_VISCO, true>(ctx, A) : launch_hex8_eas_any<MAT_VISCO, false>(ctx, A);
    if (ctx->matkind == B200S_MAT_PLASTIC_NLNLOG)
      return stiff ? launch_hex8_eas_any<MAT_PLASTIC, true>(ctx, A) : launch_hex8_eas_any<MAT_PLASTIC, false>(ctx, A);
    B200S_DISPATCH_MATERIAL(launch_hex8_eas_any);
  }
  int set_tables_hex8_eas(b200s_ctx* ctx, const Hex8Const& h)
  {
    CK(cudaMemcpyToSymbolAsync(c_hex8, &h, sizeof h, 0, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
  }
  int eas_centroid(b200s_ctx* ctx)
  {
    k_eas_centroid<<<nblk(32 * ctx->n_col_eles, 128), 128, 0, ctx->stream>>>(ctx->n_col_eles, ctx->dconn, ctx->dX,
        ctx->d_eas_t0);
    ctx->launches += 1;
    return 0;
  }
  int eas_recover(b200s_ctx* ctx, int neas, double step_length)
  {
    const unsigned grid = nblk(32 * ctx->n_col_eles, 128);
    if (!grid) return 0;
    if (neas == 7)
      k_eas_recover<7><<<grid, 128, 0, ctx->stream>>>(ctx->n_col_eles, ctx->dconn, ctx->d_du_col, step_length, ctx->d_eas_hist);
    else if (neas == 9)
      k_eas_recover<9><<<grid, 128, 0, ctx->stream>>>(ctx->n_col_eles, ctx->dconn, ctx->d_du_col, step_length,
          ctx->d_eas_hist);
    else
      k_eas_recover<21><<<grid, 128, 0, ctx->stream>>>(ctx->n_col_eles, ctx->dconn, ctx->d_du_col, step_length,
          ctx->d_eas_hist);
    ctx->launches += 1;
    return 0;
  }
}  // namespace b200s_parts
#endif  // B200S_PART_EAS

#if B200S_HAS(B200S_PART_HEX8)
namespace
{
  template <int MAT, bool STIFF>
  static int launch_hex8_t(b200s_ctx* ctx, const EvalArgs& A)
  {
    if constexpr (MAT != MAT_DP && MAT != MAT_VISCO && MAT != MAT_PLASTIC)
    if (ctx->tech == B200S_TECH_FBAR)
    {
      auto kern = k_hex8_eval_assemble<MAT, STIFF, 1>;
      if (set_smem_attr(kern, hex8_fbar_smem_bytes))
        return fail(ctx, B200S_ERR_CUDA, "cudaFuncSetAttribute(smem=%zu) failed", hex8_fbar_smem_bytes);
      const int64_t n_tiles = (A.n_ele + hex8::EPW - 1) / hex8::EPW;
      int64_t grid = (n_tiles + HEX8_FBAR_WARPS - 1) / HEX8_FBAR_WARPS;
      if (grid > ctx->n_sm) grid = ctx->n_sm;
      if (grid < 1) grid = 1;
      kern<<<(unsigned)grid, HEX8_FBAR_WARPS * 32, hex8_fbar_smem_bytes, ctx->stream>>>(A);
      ctx->launches += 1;
      return 0;
    }
    // history materials exist for one formulation each (MAT_DP: KINEM linear, MAT_VISCO: nonlinear total
    // Lagrangean): one kernel instead of four
    auto kern = [&]
    {
      if constexpr (MAT == MAT_DP)
        return k_hex8_eval_assemble<MAT, STIFF, 3>;
      else if constexpr (MAT == MAT_VISCO || MAT == MAT_PLASTIC)
        return k_hex8_eval_assemble<MAT, STIFF, 0>;
      else
        return ctx->prestress == B200S_PRESTRESS_MULF
                   ? k_hex8_eval_assemble<MAT, STIFF, 2>
                   : (ctx->kinem == B200S_KINEM_LINEAR ? k_hex8_eval_assemble<MAT, STIFF, 3>
                                                       : k_hex8_eval_assemble<MAT, STIFF, 0>);
    }();
    if constexpr (MAT == MAT_COUP_NEOHOOKE && STIFF)
      if (ctx->tma_gather && ctx->prestress != B200S_PRESTRESS_MULF && ctx->kinem != B200S_KINEM_LINEAR)
        kern = k_hex8_eval_assemble<MAT, STIFF, 0, true>;
    // per device and cheap: set on every launch (a process may hold contexts on several GPUs)
    if (set_smem_attr(kern, hex8_smem_bytes))
      return fail(ctx, B200S_ERR_CUDA, "cudaFuncSetAttribute(smem=%zu) failed", hex8_smem_bytes);
    const int64_t n_tiles = (A.n_ele + hex8::EPW - 1) / hex8::EPW;
    int64_t grid = (n_tiles + HEX8_WARPS - 1) / HEX8_WARPS;
    const int64_t cap = (int64_t)ctx->n_sm * ctx->ctas_per_sm;
    if (grid > cap) grid = cap;
    if (grid < 1) grid = 1;
    kern<<<(unsigned)grid, HEX8_WARPS * 32, hex8_smem_bytes, ctx->stream>>>(A);
    ctx->launches += 1;
    return 0;
  }

}  // namespace
namespace b200s_parts
{
  int launch_hex8(b200s_ctx* ctx, const EvalArgs& A, bool stiff)
  {
    if (ctx->matkind == B200S_MAT_DRUCKER_PRAGER)
      return stiff ? launch_hex8_t<MAT_DP, true>(ctx, A) : launch_hex8_t<MAT_DP, false>(ctx, A);
    if (ctx->matkind == B200S_MAT_VISCONEOHOOKE)
      return stiff ? launch_hex8_t<MAT_VISCO, true>(ctx, A) : launch_hex8_t<MAT_VISCO, false>(ctx, A);
    if (ctx->matkind == B200S_MAT_PLASTIC_NLNLOG)
      return stiff ? launch_hex8_t<MAT_PLASTIC, true>(ctx, A) : launch_hex8_t<MAT_PLASTIC, false>(ctx, A);
    B200S_DISPATCH_MATERIAL(launch_hex8_t);
  }
  int set_tables_hex8(b200s_ctx* ctx, const Hex8Const& h)
  {
    CK(cudaMemcpyToSymbolAsync(c_hex8, &h, sizeof h, 0, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
  }
}  // namespace b200s_parts
#endif  // B200S_PART_HEX8

#if B200S_HAS(B200S_PART_TET10)
namespace
{
  template <int MAT, bool STIFF>
  static int launch_tet10_t(b200s_ctx* ctx, const EvalArgs& A)
  {
    auto kern = [&]
    {
      if constexpr (MAT == MAT_DP)
        return k_tet10_eval_assemble<MAT, STIFF, 2>;
      else if constexpr (MAT == MAT_VISCO)
        return k_tet10_eval_assemble<MAT, STIFF, 0>;
      else
        return ctx->prestress == B200S_PRESTRESS_MULF
                   ? k_tet10_eval_assemble<MAT, STIFF, 1>
                   : (ctx->kinem == B200S_KINEM_LINEAR ? k_tet10_eval_assemble<MAT, STIFF, 2>
                                                       : k_tet10_eval_assemble<MAT, STIFF, 0>);
    }();
    // per device and cheap: set on every launch (a process may hold contexts on several GPUs)
    if (set_smem_attr(kern, tet10_smem_bytes))
      return fail(ctx, B200S_ERR_CUDA, "cudaFuncSetAttribute(smem=%zu) failed", tet10_smem_bytes);
    const int64_t n_tiles = (A.n_ele + tet10::EPW - 1) / tet10::EPW;
    int64_t grid = (n_tiles + TET10_WARPS - 1) / TET10_WARPS;
    const int64_t cap = (int64_t)ctx->n_sm * ctx->ctas_per_sm;
    if (grid > cap) grid = cap;
    if (grid < 1) grid = 1;
    kern<<<(unsigned)grid, TET10_WARPS * 32, tet10_smem_bytes, ctx->stream>>>(A);
    ctx->launches += 1;
    return 0;
  }

}  // namespace
namespace b200s_parts
{
  int launch_tet10(b200s_ctx* ctx, const EvalArgs& A, bool stiff)
  {
    if (ctx->matkind == B200S_MAT_DRUCKER_PRAGER)
      return stiff ? launch_tet10_t<MAT_DP, true>(ctx, A) : launch_tet10_t<MAT_DP, false>(ctx, A);
    if (ctx->matkind == B200S_MAT_VISCONEOHOOKE)
      return stiff ? launch_tet10_t<MAT_VISCO, true>(ctx, A) : launch_tet10_t<MAT_VISCO, false>(ctx, A);
    B200S_DISPATCH_MATERIAL(launch_tet10_t);
  }
  int set_tables_tet10(b200s_ctx* ctx, const Tet10Const& h)
  {
    CK(cudaMemcpyToSymbolAsync(c_tet10, &h, sizeof h, 0, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
  }
}  // namespace b200s_parts
#endif  // B200S_PART_TET10

#if B200S_HAS(B200S_PART_HEX27)
namespace
{
  template <int MAT, bool STIFF>
  static int launch_hex27_t(b200s_ctx* ctx, const EvalArgs& A)
  {
    if (A.nn_real > 0 && (MAT == MAT_DP || MAT == MAT_VISCO || !A.tab_override))
      return fail(ctx, B200S_ERR_UNSUPPORTED, "hex20 through the hex27 kernel: displacement-based blocks of the hyperelastic kinds");
    auto kern = [&]
    {
      if constexpr (MAT == MAT_DP)
        return k_hex27_eval_assemble<MAT, STIFF, 2>;
      else if constexpr (MAT == MAT_VISCO)
        return k_hex27_eval_assemble<MAT, STIFF, 0>;
      else if (A.nn_real > 0)  // hex20 / hex18 block (MULF: hex20 only, its 27 Gauss points are the kernel's)
        return ctx->prestress == B200S_PRESTRESS_MULF
                   ? k_hex27_eval_assemble<MAT, STIFF, 1, true>
                   : (ctx->kinem == B200S_KINEM_LINEAR ? k_hex27_eval_assemble<MAT, STIFF, 2, true>
                                                       : k_hex27_eval_assemble<MAT, STIFF, 0, true>);
      else
        return ctx->prestress == B200S_PRESTRESS_MULF
                   ? k_hex27_eval_assemble<MAT, STIFF, 1>
                   : (ctx->kinem == B200S_KINEM_LINEAR ? k_hex27_eval_assemble<MAT, STIFF, 2>
                                                       : k_hex27_eval_assemble<MAT, STIFF, 0>);
    }();
    // per device and cheap: set on every launch (a process may hold contexts on several GPUs)
    if (set_smem_attr(kern, hex27_smem_bytes))
      return fail(ctx, B200S_ERR_CUDA, "cudaFuncSetAttribute(smem=%zu) failed", hex27_smem_bytes);
    int64_t grid = (A.n_ele + HEX27_PAIRS - 1) / HEX27_PAIRS;
    if (grid > ctx->n_sm) grid = ctx->n_sm;
    if (grid < 1) grid = 1;
    kern<<<(unsigned)grid, HEX27_PAIRS * 64, hex27_smem_bytes, ctx->stream>>>(A);
    ctx->launches += 1;
    return 0;
  }

}  // namespace
namespace b200s_parts
{
  int launch_hex27(b200s_ctx* ctx, const EvalArgs& A, bool stiff)
  {
    if (ctx->matkind == B200S_MAT_DRUCKER_PRAGER)
      return stiff ? launch_hex27_t<MAT_DP, true>(ctx, A) : launch_hex27_t<MAT_DP, false>(ctx, A);
    if (ctx->matkind == B200S_MAT_VISCONEOHOOKE)
      return stiff ? launch_hex27_t<MAT_VISCO, true>(ctx, A) : launch_hex27_t<MAT_VISCO, false>(ctx, A);
    B200S_DISPATCH_MATERIAL(launch_hex27_t);
  }
  int set_tables_hex27(b200s_ctx* ctx, const Hex27Const& h)
  {
    CK(cudaMemcpyToSymbolAsync(c_hex27, &h, sizeof h, 0, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
  }
}  // namespace b200s_parts
#endif  // B200S_PART_HEX27

#if B200S_HAS(B200S_PART_CORE)
namespace
{
  // blocks evaluated by the table-driven general kernel (generic_kernel.cuh): tangents without major symmetry, and
  // the cell types without a specialised kernel (tet4, hex20)
  template <int MAT>
  static int launch_generic_mat(b200s_ctx* ctx, const GenericArgs& Ga, unsigned grid)
  {
    switch (ctx->nn)
    {
      case 5:
        k_generic_eval_assemble<MAT, 5><<<grid, 64, 0, ctx->stream>>>(Ga);
        break;
      case 6:
        k_generic_eval_assemble<MAT, 6><<<grid, 64, 0, ctx->stream>>>(Ga);
        break;
      case 4:
        k_generic_eval_assemble<MAT, 4><<<grid, 64, 0, ctx->stream>>>(Ga);
        break;
      case 8:
        k_generic_eval_assemble<MAT, 8><<<grid, 64, 0, ctx->stream>>>(Ga);
        break;
      case 10:
        k_generic_eval_assemble<MAT, 10><<<grid, 64, 0, ctx->stream>>>(Ga);
        break;
      case 18:
        k_generic_eval_assemble<MAT, 18><<<grid, 64, 0, ctx->stream>>>(Ga);
        break;
      case 20:
        k_generic_eval_assemble<MAT, 20><<<grid, 64, 0, ctx->stream>>>(Ga);
        break;
      case 27:
        k_generic_eval_assemble<MAT, 27><<<grid, 64, 0, ctx->stream>>>(Ga);
        break;
      default:
        return fail(ctx, B200S_ERR_UNSUPPORTED, "cell type %d has no device kernel", ctx->celltype);
    }
    ctx->launches += 1;
    return 0;
  }
  static int launch_generic(b200s_ctx* ctx, const EvalArgs& A, bool stiff)
  {
    GenericArgs Ga;
    Ga.A = A;
    Ga.dN_gp = ctx->d_dN_gp, Ga.dN_node = ctx->d_dN_node, Ga.w = ctx->d_w;
    Ga.nn = ctx->nn, Ga.ngp = ctx->ngp, Ga.pos_stride = ctx->pos_stride, Ga.stiff = stiff;
    // TECH fbar (hex8, pyramid5; with PRESTRESS_TECH mulf: MulfFBarFormulation)
    Ga.dN_cen = ctx->tech == B200S_TECH_FBAR ? ctx->d_dN_cen : nullptr;
    Ga.hist_pts = mulf_points(ctx);
    // TECH shell_ans (hex8): sampling-point tables
    const bool ans = ctx->tech == B200S_TECH_SHELL_ANS || ctx->tech == B200S_TECH_SHELL_EAS_ANS;
    Ga.dN_ans = ans ? ctx->d_dN_ans : nullptr;
    Ga.xi_gp = ctx->d_xi_gp;
    if (ans && (ctx->celltype != B200S_HEX8 || !ctx->d_dN_ans))
      return fail(ctx, B200S_ERR_UNSUPPORTED, "TECH shell_ans / shell_eas_ans on the device: hex8 blocks");
    // TECH shell_eas_ans: the seven EAS modes on top (history and centroid table of the block's elements)
    Ga.eas_hist = nullptr, Ga.eas_t0 = nullptr;
    if (ctx->tech == B200S_TECH_SHELL_EAS_ANS)
    {
      static_assert(ANS_EAS_HIST == hex8eas::hist_size(ANS_NEAS) && ANS_EAS_T0_STRIDE == hex8eas::T0_STRIDE, "EAS history layout");
      if (!ctx->d_eas_hist || !ctx->d_eas_t0) return fail(ctx, B200S_ERR_INVALID, "EAS history missing");
      Ga.eas_hist = ctx->d_eas_hist + A.e_begin * ANS_EAS_HIST, Ga.eas_t0 = ctx->d_eas_t0 + A.e_begin * ANS_EAS_T0_STRIDE;
    }
    const unsigned grid = nblk(A.n_ele * ctx->nn, 64);
    if (!grid) return 0;
    if (ctx->tech != B200S_TECH_NONE && ctx->tech != B200S_TECH_FBAR && !ans)
      return fail(ctx, B200S_ERR_UNSUPPORTED,
          "the general kernel evaluates displacement-based, F-bar, shell_ans and shell_eas_ans blocks only (no plain EAS)");
    switch (ctx->matkind)
    {
      case B200S_MAT_COUP_NEOHOOKE:
        return launch_generic_mat<MAT_COUP_NEOHOOKE>(ctx, Ga, grid);
      case B200S_MAT_ISO_VOL:
        return launch_generic_mat<MAT_ISO_VOL>(ctx, Ga, grid);
      case B200S_MAT_SVK:
        return launch_generic_mat<MAT_SVK>(ctx, Ga, grid);
      case B200S_MAT_ELASTHYPER:
        return launch_generic_mat<MAT_HYPER>(ctx, Ga, grid);
      case B200S_MAT_DRUCKER_PRAGER:
        return launch_generic_mat<MAT_DP>(ctx, Ga, grid);
      case B200S_MAT_VISCONEOHOOKE:
        return launch_generic_mat<MAT_VISCO>(ctx, Ga, grid);
      case B200S_MAT_PLASTIC_NLNLOG:
        return launch_generic_mat<MAT_PLASTIC>(ctx, Ga, grid);
    }
    return fail(ctx, B200S_ERR_UNSUPPORTED, "general kernel: material kind %d", ctx->matkind);
  }

  // cell types that only the table-driven general kernel hosts
  static bool is_general_cell(int ct)
  {
    return ct == B200S_TET4 || ct == B200S_HEX20 || ct == B200S_WEDGE6 || ct == B200S_PYRAMID5 || ct == B200S_HEX18;
  }

  // hex20 / hex18 displacement-based blocks with a major-symmetric tangent run through the hex27 kernel (shape tables padded
  // to 27 nodes -- hex18 also to 27 Gauss points, the padding points with weight zero --, padding rows / columns masked in
  // the scatter map): 7 x / 4 x the rate of the general kernel, whose row threads repeat the Gauss-point work nn times.  Option general_kernel = 1 keeps the general kernel (A/B, tests).
  static bool hex20_as_hex27(const b200s_ctx* ctx)
  {
    const bool hyper = ctx->matkind == B200S_MAT_COUP_NEOHOOKE || ctx->matkind == B200S_MAT_ISO_VOL ||
                       ctx->matkind == B200S_MAT_SVK || ctx->matkind == B200S_MAT_ELASTHYPER;
    // MULF histories are indexed by the cell's own Gauss points: hex20 has the kernel's 27, hex18 does not
    const bool prestress_ok = ctx->prestress == B200S_PRESTRESS_NONE ||
                              (ctx->prestress == B200S_PRESTRESS_MULF && ctx->celltype == B200S_HEX20 && ctx->d_gp_hist);
    return (ctx->celltype == B200S_HEX20 || ctx->celltype == B200S_HEX18) && !ctx->general_tangent && !ctx->general_kernel && ctx->tech == B200S_TECH_NONE &&
           prestress_ok && hyper && ctx->d_conn27 && ctx->d_pos27 && ctx->d_tab20 && ctx->d_dN_node;
  }

  template <bool STIFF>
  static int launch_eval(b200s_ctx* ctx, const EvalArgs& A)
  {
    if (ctx->matkind < 0) return fail(ctx, B200S_ERR_INVALID, "material not set");
    if (hex20_as_hex27(ctx))
    {
      EvalArgs B = A;
      B.conn = ctx->d_conn27 + A.e_begin * 27;
      B.pos = ctx->d_pos27 + A.e_begin * hex27::POS_STRIDE;
      B.tab_override = ctx->d_tab20, B.nn_real = ctx->nn;
      k_nodal_jacobian_check<<<nblk(A.n_ele, 128), 128, 0, ctx->stream>>>(ctx->nn, A.n_ele, A.conn, A.X, ctx->d_dN_node, A.ele_id,
          A.err);
      ctx->launches += 1;
      return b200s_parts::launch_hex27(ctx, B, STIFF);
    }
    if (ctx->general_tangent || ctx->general_kernel || is_general_cell(ctx->celltype) ||
        (ctx->tech == B200S_TECH_FBAR && ctx->prestress == B200S_PRESTRESS_MULF) || ctx->tech == B200S_TECH_SHELL_ANS ||
        ctx->tech == B200S_TECH_SHELL_EAS_ANS || (ctx->matkind == B200S_MAT_PLASTIC_NLNLOG && ctx->celltype != B200S_HEX8))  // fast kernels: hex8 (+ EAS) only
      return launch_generic(ctx, A, STIFF);
    if (ctx->tech != B200S_TECH_NONE && ctx->celltype != B200S_HEX8)
      return fail(ctx, B200S_ERR_UNSUPPORTED, "element technology %d needs hex8", ctx->tech);
    switch (ctx->celltype)
    {
      case B200S_HEX27:
        return b200s_parts::launch_hex27(ctx, A, STIFF);
      case B200S_HEX8:
        if (eas_neas(ctx->tech)) return b200s_parts::launch_hex8_eas(ctx, A, STIFF);
        return b200s_parts::launch_hex8(ctx, A, STIFF);
      case B200S_TET10:
        return b200s_parts::launch_tet10(ctx, A, STIFF);
    }
    return fail(ctx, B200S_ERR_UNSUPPORTED, "cell type %d has no device kernel", ctx->celltype);
  }

}  // namespace

extern "C"
{
  const char* b200s_version(void) { return "4c_b200 solid 0.1 (sm_100a, fp64)"; }

  int b200s_create(int device, int rank, int nranks, b200s_ctx** out)
  {
    if (!out) return B200S_ERR_INVALID;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0 || device >= ndev)
      return B200S_ERR_CUDA;  // no CUDA device: the product has no CPU fallback
    b200s_ctx* ctx = new b200s_ctx;
    ctx->device = device, ctx->rank = rank, ctx->nranks = nranks;
    cudaSetDevice(device);
    cudaDeviceGetAttribute(&ctx->n_sm, cudaDevAttrMultiProcessorCount, device);
    if (cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) != cudaSuccess ||
        (ctx->stream = ctx->own_stream) == nullptr ||
        cudaMalloc((void**)&ctx->d_err, sizeof(ErrWord)) != cudaSuccess ||
        cudaMallocHost((void**)&ctx->h_err, 2 * sizeof(ErrWord)) != cudaSuccess)
    {
      delete ctx;
      return B200S_ERR_CUDA;
    }
    ctx->h_err_init = ctx->h_err + 1;
    for (ErrWord* w : {ctx->h_err, ctx->h_err_init})
    {
      w->n_neg_jac = 0, w->first_bad = LLONG_MAX;
      w->min_detJ = 1e300, w->min_detF = 1e300, w->n_nonfinite = 0;
    }
    for (auto& e : ctx->ev) cudaEventCreate(&e);
    for (auto& e : ctx->kev) cudaEventCreate(&e);
    for (auto& e : ctx->ev_la) cudaEventCreate(&e);
    cudaStreamCreateWithFlags(&ctx->zero_stream, cudaStreamNonBlocking);
    cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking);
    cudaStreamCreateWithFlags(&ctx->halo_stream, cudaStreamNonBlocking);
    for (auto& e : ctx->ev_phase) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&ctx->ev_halo_done, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&ctx->ev_copy_done, cudaEventDisableTiming);
    for (auto& e : ctx->ev_zero_done) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&ctx->ev_eval_start, cudaEventDisableTiming);
    *out = ctx;
    return B200S_OK;
  }

  void b200s_destroy(b200s_ctx* ctx)
  {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->copy_stream) cudaStreamSynchronize(ctx->copy_stream), cudaStreamDestroy(ctx->copy_stream);
    if (ctx->halo_stream) cudaStreamSynchronize(ctx->halo_stream), cudaStreamDestroy(ctx->halo_stream);
    for (auto& e : ctx->ev_phase) cudaEventDestroy(e);
    cudaEventDestroy(ctx->ev_fork), cudaEventDestroy(ctx->ev_halo_done), cudaEventDestroy(ctx->ev_copy_done);
    delete ctx->sym;
    ctx->sym = nullptr;
    if (ctx->primary) detach_shared(ctx);  // the primary owns graph, values, vectors, stream and communicator
    if (ctx->comm && ctx->nccl.CommDestroy) ctx->nccl.CommDestroy(ctx->comm);
    dfree(ctx->dX), dfree(ctx->dconn), dfree(ctx->d_per_ele), dfree(ctx->d_rowptr);
    dfree(ctx->d_colind), dfree(ctx->d_pos), dfree(ctx->d_node_row), dfree(ctx->d_ele_id);
    if (ctx->zero_stream) cudaStreamSynchronize(ctx->zero_stream);
    dfree(ctx->d_node_colbase), dfree(ctx->d_vals_buf[0]), dfree(ctx->d_vals_buf[1]), dfree(ctx->d_f);
    dfree(ctx->d_zero);
    ctx->d_vals = nullptr;
    for (auto& e : ctx->ev_zero_done) cudaEventDestroy(e);
    cudaEventDestroy(ctx->ev_eval_start);
    if (ctx->zero_stream) cudaStreamDestroy(ctx->zero_stream);
    dfree(ctx->d_u_row), dfree(ctx->d_u_col), dfree(ctx->d_send_lid), dfree(ctx->d_recv_lid);
    dfree(ctx->d_sendbuf), dfree(ctx->d_recvbuf), dfree(ctx->d_err);
    dfree(ctx->d_gp_hist), dfree(ctx->d_eas_hist), dfree(ctx->d_du_row), dfree(ctx->d_du_col);
    dfree(ctx->d_eas_t0), dfree(ctx->d_fibers), dfree(ctx->d_blkcol), dfree(ctx->d_hist_old);
    dfree(ctx->d_row_diagcol);
    dfree(ctx->d_xi_gp), dfree(ctx->d_dN_node);
    dfree(ctx->d_mat_ext), dfree(ctx->d_mat_hist);
    b200s_clear_neumann(ctx);
    dfree(ctx->d_face_N), dfree(ctx->d_face_d0), dfree(ctx->d_face_d1), dfree(ctx->d_face_w);
    dfree(ctx->d_N_gp), dfree(ctx->d_fext), dfree(ctx->d_nm_row), dfree(ctx->d_nm_col), dfree(ctx->d_nm_flag);
    dfree(ctx->d_exp_pos), dfree(ctx->d_exp_fpos), dfree(ctx->d_exp_buf), dfree(ctx->d_exp_fbuf);
    dfree(ctx->d_hyper), dfree(ctx->d_dN_gp), dfree(ctx->d_w), dfree(ctx->d_dN_cen), dfree(ctx->d_N_mass), dfree(ctx->d_dN_pts), dfree(ctx->d_dN_ans);
    dfree(ctx->d_conn27), dfree(ctx->d_pos27);
    if (ctx->d_tab20) cudaFree(ctx->d_tab20);
    ctx->d_tab20 = nullptr;
    dfree(ctx->d_dN_mass), dfree(ctx->d_w_mass), dfree(ctx->d_ele_is_row), dfree(ctx->d_mass);
    dfree(ctx->d_f_inertia), dfree(ctx->d_acc_row), dfree(ctx->d_acc_col), dfree(ctx->d_energy);
    dfree(ctx->d_xmat), dfree(ctx->d_la_x), dfree(ctx->d_la_y), dfree(ctx->d_toggle);
    for (auto& v : ctx->d_cg) dfree(v);
    dfree(ctx->d_cg_s), dfree(ctx->d_cg_hist);
    for (auto& e : ctx->ev_la) cudaEventDestroy(e);
    if (ctx->h_err) cudaFreeHost(ctx->h_err);
    for (auto& e : ctx->ev) cudaEventDestroy(e);
    for (auto& e : ctx->kev) cudaEventDestroy(e);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
  }

  const char* b200s_last_error(const b200s_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

  int b200s_set_option(b200s_ctx* ctx, const char* key, double value)
  {
    if (!ctx || !key) return B200S_ERR_INVALID;
    if (!strcmp(key, "check_finite"))
      ctx->check_finite = value != 0;
    else if (!strcmp(key, "ctas_per_sm"))
      ctx->ctas_per_sm = (int)value;
    else if (!strcmp(key, "scatter_mode"))
      ctx->scatter_mode = (int)value;
    else if (!strcmp(key, "reorder"))  // takes effect at the next set_mesh: 0 caller's order, 1 Z-order curve, 2 Hilbert curve
      ctx->reorder = (int)value;
    else if (!strcmp(key, "double_buffer"))  // takes effect at the next set_graph / build_graph
      ctx->double_buffer = (int)value;  // 0 off, 1 memset on a side stream, 2 copy-engine zeroing on a side stream
    else if (!strcmp(key, "pipeline_phases"))  // takes effect at the next set_graph / build_graph
      ctx->pipeline_phases = (int)value;
    else if (!strcmp(key, "overlap_halo"))
      ctx->overlap_halo = value != 0;
    else if (!strcmp(key, "owned_rows"))  // row-element mode, before set_graph: dof rows the rank owns
      ctx->n_rows_owned = (int64_t)value;
    else if (!strcmp(key, "owned_nnz"))  // ... and the non-zeros of those rows
      ctx->nnz_owned = (int64_t)value;
    else if (!strcmp(key, "general_kernel"))  // A/B and tests: route every block through the table-driven general kernel
      ctx->general_kernel = value != 0;
    else if (!strcmp(key, "tma_gather"))
      ctx->tma_gather = value != 0;
    else if (!strcmp(key, "eas_mma"))
      ctx->eas_mma = value != 0;
    else if (!strcmp(key, "segments"))  // takes effect at the next set_graph / build_graph
      ctx->segments = value != 0;
    else if (!strcmp(key, "symmetric_transfer"))
      ctx->symmetric_transfer = value != 0;
    else if (!strcmp(key, "host_threads"))  // takes effect when the transfer tables are (re)built
    {
      ctx->host_threads = (int)value;
      delete ctx->sym;
      ctx->sym = nullptr;
    }
    else if (!strcmp(key, "nt_stores"))
      ctx->nt_stores = value != 0;
    else
      return fail(ctx, B200S_ERR_INVALID, "unknown option %s", key);
    return B200S_OK;
  }

  int b200s_set_mesh(b200s_ctx* ctx, int celltype, int64_t n_col_nodes, const double* X,
      int64_t n_col_eles, const int32_t* conn)
  {
    // a processor may hold no element (empty column-element loop) or nothing at all
    if (!ctx || n_col_nodes < 0 || n_col_eles < 0 || (n_col_nodes > 0 && !X) || (n_col_eles > 0 && (!conn || n_col_nodes == 0)))
      return B200S_ERR_INVALID;
    cudaSetDevice(ctx->device);
    if (ctx->primary) detach_shared(ctx);  // a new mesh: attach again afterwards
    if (celltype != B200S_HEX8 && celltype != B200S_TET10 && celltype != B200S_HEX27 && !is_general_cell(celltype))
      return fail(ctx, B200S_ERR_UNSUPPORTED, "cell type %d is not on the device path", celltype);
    if (n_col_nodes > INT32_MAX / 3) return fail(ctx, B200S_ERR_INVALID, "too many column nodes");
    const CellTables T = make_cell_tables(celltype);
    ctx->celltype = celltype, ctx->nn = T.nn;
    ctx->n_col_nodes = n_col_nodes, ctx->n_col_eles = n_col_eles;
    if (to_device(ctx, &ctx->dX, X, 3 * n_col_nodes)) return B200S_ERR_CUDA;
    if (to_device(ctx, &ctx->dconn, conn, n_col_eles * T.nn)) return B200S_ERR_CUDA;
    if (to_device(ctx, &ctx->d_u_col, (const double*)nullptr, 3 * n_col_nodes)) return B200S_ERR_CUDA;
    CK(cudaMemsetAsync(ctx->d_u_col, 0, sizeof(double) * 3 * n_col_nodes, ctx->stream));
    ctx->ngp = T.ngp, ctx->ngp_mass = T.ngp_mass;
    if (to_device(ctx, &ctx->d_dN_gp, T.dN_gp.data(), (int64_t)T.dN_gp.size()) ||
        to_device(ctx, &ctx->d_w, T.w.data(), (int64_t)T.w.size()) ||
        to_device(ctx, &ctx->d_dN_cen, T.dN_cen.data(), (int64_t)T.dN_cen.size()) ||
        (!T.dN_ans.empty() && to_device(ctx, &ctx->d_dN_ans, T.dN_ans.data(), (int64_t)T.dN_ans.size())) ||
        to_device(ctx, &ctx->d_dN_pts, (const double*)nullptr, (int64_t)(T.dN_gp.size() + T.dN_cen.size())) ||
        to_device(ctx, &ctx->d_dN_node, T.dN_node.data(), (int64_t)T.dN_node.size()) ||
        to_device(ctx, &ctx->d_xi_gp, T.xi_gp.data(), (int64_t)T.xi_gp.size()) ||
        to_device(ctx, &ctx->d_N_mass, T.N_mass.data(), (int64_t)T.N_mass.size()) ||
        to_device(ctx, &ctx->d_dN_mass, T.dN_mass.data(), (int64_t)T.dN_mass.size()) ||
        to_device(ctx, &ctx->d_w_mass, T.w_mass.data(), (int64_t)T.w_mass.size()) ||
        to_device(ctx, &ctx->d_N_gp, T.N_gp.data(), (int64_t)T.N_gp.size()))
      return B200S_ERR_CUDA;
    CK(cudaMemcpyAsync(ctx->d_dN_pts, ctx->d_dN_gp, 8 * T.dN_gp.size(), cudaMemcpyDeviceToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->d_dN_pts + T.dN_gp.size(), ctx->d_dN_cen, 8 * T.dN_cen.size(), cudaMemcpyDeviceToDevice, ctx->stream));
    const FaceTables FT = make_face_tables(celltype);
    ctx->face_nn = FT.nn, ctx->face_ngp = FT.ngp;
    if (to_device(ctx, &ctx->d_face_N, FT.N.data(), (int64_t)FT.N.size()) ||
        to_device(ctx, &ctx->d_face_d0, FT.d0.data(), (int64_t)FT.d0.size()) ||
        to_device(ctx, &ctx->d_face_d1, FT.d1.data(), (int64_t)FT.d1.size()) ||
        to_device(ctx, &ctx->d_face_w, FT.w.data(), (int64_t)FT.w.size()))
      return B200S_ERR_CUDA;
    b200s_clear_neumann(ctx);  // conditions refer to the previous mesh
    dfree(ctx->d_nm_col);
    CK(cudaStreamSynchronize(ctx->stream));  // T is a local
    dfree(ctx->d_ele_is_row), dfree(ctx->d_acc_col), dfree(ctx->d_gp_hist), dfree(ctx->d_eas_hist);
    dfree(ctx->d_du_col), dfree(ctx->d_fibers), dfree(ctx->d_hist_old);
    // per-element data of the previous mesh (caller's numbering of THAT mesh): the caller sets it again
    dfree(ctx->d_per_ele);
    if (celltype == B200S_HEX8)
    {
      Hex8Const h;
      memcpy(h.dN_gp, T.dN_gp.data(), sizeof h.dN_gp);
      memcpy(h.dN_node, T.dN_node.data(), sizeof h.dN_node);
      memcpy(h.w, T.w.data(), sizeof h.w);
      memcpy(h.xi_gp, T.xi_gp.data(), sizeof h.xi_gp);
      if (b200s_parts::set_tables_hex8(ctx, h) || b200s_parts::set_tables_hex8_eas(ctx, h)) return B200S_ERR_CUDA;
    }
    else if (celltype == B200S_HEX27)
    {
      std::vector<Hex27Const> hbuf(1);  // 18 KB: not on the stack
      Hex27Const& h = hbuf[0];
      memset(&h, 0, sizeof h);
      memcpy(h.dN_gp, T.dN_gp.data(), sizeof h.dN_gp);
      memcpy(h.w, T.w.data(), 27 * sizeof(double));
      const double(*hc)[3] = detail::hex_nodes();
      for (int k = 0; k < 27; ++k)
      {
        const int a = (int)hc[k][0] + 1, b = (int)hc[k][1] + 1, c = (int)hc[k][2] + 1;
        h.it.abc[k] = a | (b << 2) | (c << 4);
        h.it.node_of[a + 3 * b + 9 * c] = k;
      }
      if (b200s_parts::set_tables_hex27(ctx, h)) return B200S_ERR_CUDA;
    }
    else if (celltype == B200S_HEX20 || celltype == B200S_HEX18)
    {
      // hex20 / hex18 blocks through the hex27 kernel: the cell's shape derivatives at its Gauss points in the hex27 table
      // layout, zero for the padding nodes (hex20: the same hex_27point rule, 7 padding nodes; hex18: hex_18point, 9
      // padding nodes and 9 padding points = copies of point 0 with weight zero, which contribute exactly nothing); the
      // integer tables are hex27's (only the native hex27 launches use them)
      std::vector<Hex27Const> hbuf(1);
      Hex27Const& h = hbuf[0];
      memset(&h, 0, sizeof h);
      pad_tables_to_hex27(T, h.dN_gp, h.w);
      const double(*hc)[3] = detail::hex_nodes();
      for (int k = 0; k < 27; ++k)
      {
        const int a = (int)hc[k][0] + 1, b = (int)hc[k][1] + 1, c = (int)hc[k][2] + 1;
        h.it.abc[k] = a | (b << 2) | (c << 4);
        h.it.node_of[a + 3 * b + 9 * c] = k;
      }
      Hex27Const* d = nullptr;
      if (to_device(ctx, &d, &h, 1)) return B200S_ERR_CUDA;
      CK(cudaStreamSynchronize(ctx->stream));  // h leaves scope
      if (ctx->d_tab20) cudaFree(ctx->d_tab20);
      ctx->d_tab20 = d;
    }
    else if (celltype == B200S_TET10)
    {
      Tet10Const h;
      memcpy(h.dN_gp, T.dN_gp.data(), sizeof h.dN_gp);
      memcpy(h.dN_node, T.dN_node.data(), sizeof h.dN_node);
      memcpy(h.w, T.w.data(), sizeof h.w);
      if (b200s_parts::set_tables_tet10(ctx, h)) return B200S_ERR_CUDA;
    }
    dfree(ctx->d_ele_id);
    if (ctx->reorder && n_col_eles > 1)
    {
      if (n_col_eles > INT32_MAX) return fail(ctx, B200S_ERR_INVALID, "too many column elements");
      cudaStream_t st = ctx->stream;
      double hbox[6] = {1e300, 1e300, 1e300, -1e300, -1e300, -1e300};
      double* d_box = nullptr;
      unsigned long long *k0 = nullptr, *k1 = nullptr;
      int32_t *i0 = nullptr, *conn2 = nullptr;
      void* tmp = nullptr;
      CK(cudaMalloc((void**)&d_box, sizeof hbox));
      CK(cudaMemcpyAsync(d_box, hbox, sizeof hbox, cudaMemcpyHostToDevice, st));
      CK(cudaMalloc((void**)&k0, 8 * (size_t)n_col_eles));
      CK(cudaMalloc((void**)&k1, 8 * (size_t)n_col_eles));
      CK(cudaMalloc((void**)&i0, 4 * (size_t)n_col_eles));
      CK(cudaMalloc((void**)&ctx->d_ele_id, 4 * (size_t)n_col_eles));
      CK(cudaMalloc((void**)&conn2, 4 * (size_t)n_col_eles * T.nn));
      k_bbox<<<ctx->n_sm * 4, 256, 0, st>>>(n_col_nodes, ctx->dX, d_box);
      k_morton_keys<<<nblk(n_col_eles), 256, 0, st>>>(T.nn, n_col_eles, ctx->dconn, ctx->dX, d_box, k0, i0, ctx->reorder == 2);
      size_t tb = 0;
      cub::DeviceRadixSort::SortPairs(nullptr, tb, k0, k1, i0, ctx->d_ele_id, n_col_eles, 0, 63, st);
      CK(cudaMalloc(&tmp, tb));
      cub::DeviceRadixSort::SortPairs(tmp, tb, k0, k1, i0, ctx->d_ele_id, n_col_eles, 0, 63, st);
      k_permute_conn<<<nblk(n_col_eles * T.nn), 256, 0, st>>>(T.nn, n_col_eles, ctx->dconn, ctx->d_ele_id, conn2);
      ctx->launches += 3;
      CK(cudaStreamSynchronize(st));
      CK(cudaGetLastError());
      cudaFree(ctx->dconn);
      ctx->dconn = conn2;
      cudaFree(d_box), cudaFree(k0), cudaFree(k1), cudaFree(i0), cudaFree(tmp);
    }
    CK(cudaStreamSynchronize(ctx->stream));
    // histories are indexed by the DEVICE element index: build them only now that dconn / d_ele_id are in
    // their final (locality) order -- also when set_element_properties / set_material came first
    if (ctx->prestress == B200S_PRESTRESS_MULF && mulf_allocate(ctx)) return B200S_ERR_CUDA;
    if (eas_neas(ctx->tech) && eas_allocate(ctx)) return B200S_ERR_CUDA;
    if (material_history_allocate(ctx)) return B200S_ERR_CUDA;
    // graph-dependent data is stale now (build_graph / set_graph must follow)
    dfree(ctx->d_pos), dfree(ctx->d_rowptr), dfree(ctx->d_colind), dfree(ctx->d_conn27), dfree(ctx->d_pos27);
    ctx->nnz = 0, ctx->n_rows = 0;
    return B200S_OK;
  }

  int b200s_set_element_properties(b200s_ctx* ctx, int kinem_type, int element_technology,
      int prestress_technology)
  {
    if (!ctx) return B200S_ERR_INVALID;
    if (kinem_type != B200S_KINEM_NONLINEAR_TOTLAG && kinem_type != B200S_KINEM_LINEAR)
      return fail(ctx, B200S_ERR_UNSUPPORTED, "kinematic type %d has no device kernel", kinem_type);
    // linear kinematics: displacement based, or EAS (EASFormulation<hex8, h8_9 | h8_21, linear>,
    // 4C_solid_ele_factory.hpp:71-74); F-bar / MULF have no linear variant in the reference
    const bool eas_tech = element_technology == B200S_TECH_EAS_MILD || element_technology == B200S_TECH_EAS_FULL;  // (not shell_eas)
    if (kinem_type == B200S_KINEM_LINEAR &&
        ((element_technology != B200S_TECH_NONE && !eas_tech) || prestress_technology != B200S_PRESTRESS_NONE))
      return fail(ctx, B200S_ERR_UNSUPPORTED, "KINEM linear with F-bar or prestress has no device kernel");
    if (prestress_technology != B200S_PRESTRESS_NONE && prestress_technology != B200S_PRESTRESS_MULF)
      return fail(ctx, B200S_ERR_UNSUPPORTED, "prestress technology %d has no device kernel",
          prestress_technology);
    // MulfFBarFormulation (4C_solid_ele_calc_mulf_fbar.hpp: hex8, pyramid5) runs through the general kernel; MULF with
    // EAS does not exist in the reference (4C_solid_ele_factory_lib.hpp)
    if (prestress_technology == B200S_PRESTRESS_MULF && element_technology != B200S_TECH_NONE &&
        element_technology != B200S_TECH_FBAR)
      return fail(ctx, B200S_ERR_UNSUPPORTED, "MULF prestress with element technology %d has no device kernel",
          element_technology);
    if (element_technology < B200S_TECH_NONE || element_technology > B200S_TECH_SHELL_EAS_ANS)
      return fail(ctx, B200S_ERR_UNSUPPORTED, "element technology %d has no device kernel",
          element_technology);
    // the reference offers F-bar and EAS for hex8 only among the device cell types
    // (4C_solid_ele_factory_lib.hpp)
    if (element_technology != B200S_TECH_NONE && ctx->celltype >= 0 && ctx->celltype != B200S_HEX8 &&
        !(element_technology == B200S_TECH_FBAR && ctx->celltype == B200S_PYRAMID5))
      return fail(ctx, B200S_ERR_UNSUPPORTED,
          "F-bar is a hex8 / pyramid5 technology, EAS a hex8 technology, shell_ans on the device a hex8 technology (the "
          "wedge6 solid shell stays on the CPU path)");
    // ShellANSFormulation exists for KINEM nonlinear without prestress (4C_solid_ele_factory.cpp)
    if ((element_technology == B200S_TECH_SHELL_ANS || element_technology == B200S_TECH_SHELL_EAS ||
            element_technology == B200S_TECH_SHELL_EAS_ANS) &&
        (kinem_type != B200S_KINEM_NONLINEAR_TOTLAG || prestress_technology != B200S_PRESTRESS_NONE))
      return fail(ctx, B200S_ERR_UNSUPPORTED, "TECH shell_ans / shell_eas / shell_eas_ans: KINEM nonlinear, no prestress");
    const int old_points = mulf_points(ctx);
    ctx->kinem = kinem_type;
    dfree(ctx->d_hist_old);
    const bool eas_fresh = eas_neas(element_technology) != eas_neas(ctx->tech) || !ctx->d_eas_hist;
    ctx->tech = element_technology;
    if (!eas_neas(ctx->tech)) dfree(ctx->d_eas_hist);
    if (eas_neas(ctx->tech) && eas_fresh && ctx->celltype >= 0)
    {
      cudaSetDevice(ctx->device);
      if (eas_allocate(ctx)) return B200S_ERR_CUDA;
    }
    const bool fresh = prestress_technology == B200S_PRESTRESS_MULF &&
                       (ctx->prestress != B200S_PRESTRESS_MULF || mulf_points(ctx) != old_points);
    ctx->prestress = prestress_technology;
    if (prestress_technology != B200S_PRESTRESS_MULF) dfree(ctx->d_gp_hist);
    if ((fresh || !ctx->d_gp_hist) && prestress_technology == B200S_PRESTRESS_MULF && ctx->celltype >= 0)
    {
      cudaSetDevice(ctx->device);
      if (mulf_allocate(ctx)) return B200S_ERR_CUDA;
    }
    return B200S_OK;
  }

  int b200s_set_material(b200s_ctx* ctx, int kind, const double* params, int nparams,
      const double* per_ele_param)
  {
    if (!ctx || !params || nparams < 2) return B200S_ERR_INVALID;
    cudaSetDevice(ctx->device);
    MatParams mp{};
    if (kind == B200S_MAT_DRUCKER_PRAGER)
    {
      if (nparams < 10 || per_ele_param)
        return fail(ctx, B200S_ERR_INVALID,
            "MAT_Struct_DruckerPrager: 10 parameters {YOUNG NUE ISOHARD TOL C ETA XI ETABAR MAXITER TANG}");
      double q[DP_NPAR];
      if (params[4] == 0.0 || params[5] == 0.0 || params[7] == 0.0)
        return fail(ctx, B200S_ERR_INVALID,
            "MAT_Struct_DruckerPrager: C, ETA and ETABAR must be non-zero (the yield function is normalised by the "
            "cohesion and the return mapping divides by the friction parameters)");
      if (!make_drucker_prager(params, q)) return fail(ctx, B200S_ERR_INVALID, "MAT_Struct_DruckerPrager: parameters rejected");
      // the non-associative consistent tangent (ETA != ETABAR) has no major symmetry: general kernel
      ctx->general_tangent = !drucker_prager_tangent_is_symmetric(params);
      for (double& v : mp.p) v = 0.0;
      if (to_device(ctx, &ctx->d_mat_ext, q, DP_NPAR)) return B200S_ERR_CUDA;
      CK(cudaStreamSynchronize(ctx->stream));  // q is a local
      mp.ext = ctx->d_mat_ext;
      ctx->matkind = kind, ctx->mp = mp;
      dfree(ctx->d_per_ele);
      return material_history_allocate(ctx) ? B200S_ERR_CUDA : B200S_OK;
    }
    ctx->general_tangent = false;
    if (kind == B200S_MAT_VISCONEOHOOKE)
    {
      if (nparams < 6 || per_ele_param)
        return fail(ctx, B200S_ERR_INVALID,
            "MAT_VISCONEOHOOKE: 6 parameters {YOUNGS_SLOW POISSON DENS YOUNGS_FAST RELAX THETA}");
      for (double& v : mp.p) v = 0.0;
      mp.hyper = nullptr, mp.fiber = nullptr, mp.ext = nullptr, mp.hist = nullptr;
      if (params[3] < params[0]) return fail(ctx, B200S_ERR_INVALID, "Wrong ratio between fast and slow Young's modulus");
      if (!(params[4] > 0.0)) return fail(ctx, B200S_ERR_INVALID, "Relaxation time tau has to be positive!");
      for (int i = 0; i < 6; ++i) ctx->visco_in[i] = params[i];
      make_visco_neohooke(ctx->visco_in, ctx->dt, mp.p);
      const bool fresh = ctx->matkind != kind || !ctx->d_mat_hist;
      ctx->matkind = kind, ctx->mp = mp;
      if (params[2] > 0.0) ctx->density = params[2];
      dfree(ctx->d_per_ele);
      return (fresh && material_history_allocate(ctx)) ? B200S_ERR_CUDA : B200S_OK;
    }
    if (kind == B200S_MAT_PLASTIC_NLNLOG)
    {
      if (nparams < 10 || per_ele_param)
        return fail(ctx, B200S_ERR_INVALID,
            "MAT_Struct_PlasticNlnLogNeoHooke: 10 parameters {YOUNG NUE DENS YIELD ISOHARD SATHARDENING HARDEXPO VISC "
            "RATE_DEPENDENCY TOL}");
      if (params[3] == 0.0)
        return fail(ctx, B200S_ERR_UNSUPPORTED,
            "You have to provide either a parameter for HARDENING_FUNC or YIELD in MAT_Struct_PlasticNlnLogNeoHooke "
            "(HARDENING_FUNC, a parsed function, stays on the CPU path)");
      for (int i = 0; i < 10; ++i) ctx->plastic_in[i] = params[i];
      double q[PLASTIC_NPAR];
      // the rate term divides by the time step size: a placeholder until b200s_set_time_step_size has been called
      if (!make_plastic_nlnlog(ctx->plastic_in, ctx->dt_set && ctx->dt > 0.0 ? ctx->dt : 1.0, q))
        return fail(ctx, B200S_ERR_INVALID, "MAT_Struct_PlasticNlnLogNeoHooke: parameters rejected");
      for (double& v : mp.p) v = 0.0;
      mp.hyper = nullptr, mp.fiber = nullptr, mp.hist = nullptr;
      if (to_device(ctx, &ctx->d_mat_ext, q, PLASTIC_NPAR)) return B200S_ERR_CUDA;
      CK(cudaStreamSynchronize(ctx->stream));  // q is a local
      mp.ext = ctx->d_mat_ext;
      const bool fresh = ctx->matkind != kind || !ctx->d_mat_hist;
      ctx->matkind = kind, ctx->mp = mp;
      if (params[2] > 0.0) ctx->density = params[2];
      dfree(ctx->d_per_ele);
      return (fresh && material_history_allocate(ctx)) ? B200S_ERR_CUDA : B200S_OK;
    }
    dfree(ctx->d_mat_hist);
    if (!make_mat_params(kind, params, mp))
      return fail(ctx, B200S_ERR_UNSUPPORTED, "material kind %d is not on the device path", kind);
    if (per_ele_param && kind != B200S_MAT_ISO_VOL)
      return fail(ctx, B200S_ERR_UNSUPPORTED, "per-element parameter only for ISO_VOL (MUE)");
    ctx->matkind = kind, ctx->mp = mp;
    dfree(ctx->d_per_ele);
    if (per_ele_param)
    {
      if (ctx->n_col_eles <= 0) return fail(ctx, B200S_ERR_INVALID, "set_mesh before per-element parameters");
      if (to_device(ctx, &ctx->d_per_ele, per_ele_param, ctx->n_col_eles)) return B200S_ERR_CUDA;
      if (ctx->d_ele_id)  // caller's element order -> device (locality) order
      {
        double* sorted = nullptr;
        CK(cudaMalloc((void**)&sorted, 8 * (size_t)ctx->n_col_eles));
        k_gather_double<<<nblk(ctx->n_col_eles), 256, 0, ctx->stream>>>(
            ctx->n_col_eles, ctx->d_per_ele, ctx->d_ele_id, sorted);
        ctx->launches += 1;
        CK(cudaStreamSynchronize(ctx->stream));
        cudaFree(ctx->d_per_ele);
        ctx->d_per_ele = sorted;
      }
      CK(cudaStreamSynchronize(ctx->stream));
    }
    return B200S_OK;
  }

  // Element segments for the graph just installed (b200s_ctx::seg_begin): stable sort of the locality
  // (Morton) order of set_mesh by {holds a ghost node, row phase of the lowest owned node}.  Needs
  // ctx->d_node_row and ctx->d_rowptr; re-orders every per-element array that already exists.
  static int order_elements_for_graph(b200s_ctx* ctx)
  {
    const int64_t n = ctx->n_col_eles;
    ctx->segmented = false, ctx->n_phases = 1;
    ctx->seg_begin = {0, 0, n};
    ctx->phase_row = {0, ctx->n_rows};
    ctx->phase_off = {0, ctx->nnz};
    if (!ctx->reorder || !ctx->segments || n <= 1 || !ctx->d_rowptr) return 0;
    cudaStream_t st = ctx->stream;
    int P = ctx->pipeline_phases < 1 ? 1 : (ctx->pipeline_phases > b200s_ctx::MAX_PHASES - 2 ? b200s_ctx::MAX_PHASES - 2
                                                                                              : ctx->pipeline_phases);
    while (P > 1 && ctx->nnz / P < (1 << 21)) --P;  // at least 16 MB of values per phase
    int64_t *d_prow = nullptr, *d_poff = nullptr, *d_seg = nullptr;
    uint32_t *k0 = nullptr, *k1 = nullptr;
    int32_t *i0 = nullptr, *perm = nullptr;
    unsigned long long* d_spread = nullptr;
    void* tmp = nullptr;
    CK(cudaMalloc((void**)&d_prow, 8 * b200s_ctx::MAX_PHASES));
    CK(cudaMalloc((void**)&d_poff, 8 * b200s_ctx::MAX_PHASES));
    CK(cudaMalloc((void**)&d_seg, 8 * (b200s_ctx::MAX_PHASES + 2)));
    CK(cudaMalloc((void**)&d_spread, 8));
    CK(cudaMalloc((void**)&k0, 4 * (size_t)n));
    CK(cudaMalloc((void**)&k1, 4 * (size_t)n));
    CK(cudaMalloc((void**)&i0, 4 * (size_t)n));
    CK(cudaMalloc((void**)&perm, 4 * (size_t)n));
    for (int attempt = 0; attempt < 2; ++attempt)
    {
      unsigned long long spread = 0;
      k_phase_rows<<<1, b200s_ctx::MAX_PHASES, 0, st>>>(P, ctx->n_rows, ctx->d_rowptr, d_prow, d_poff);
      CK(cudaMemsetAsync(d_spread, 0, 8, st));
      k_segment_keys<<<nblk(n), 256, 0, st>>>(ctx->nn, n, ctx->dconn, ctx->d_node_row, P, d_prow, k0, i0, d_spread, io_rows(ctx));
      ctx->launches += 2;
      CK(cudaMemcpyAsync(&spread, d_spread, 8, cudaMemcpyDeviceToHost, st));
      CK(cudaStreamSynchronize(st));
      // rows numbered without spatial locality (e.g. randomly renumbered nodes): row phases would scatter
      // every phase over the whole mesh and cost the L2 locality of the Morton order -- one phase then
      if (P > 1 && spread > (unsigned long long)(n / 8))
      {
        P = 1;
        continue;
      }
      break;
    }
    size_t tb = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tb, k0, k1, i0, perm, n, 0, 8, st);
    CK(cudaMalloc(&tmp, tb));
    cub::DeviceRadixSort::SortPairs(tmp, tb, k0, k1, i0, perm, n, 0, 8, st);  // stable: Morton order inside a segment
    k_segment_bounds<<<1, 128, 0, st>>>(P + 1, n, k1, d_seg);
    ctx->launches += 1;
    ctx->n_phases = P;
    ctx->seg_begin.assign((size_t)P + 2, 0);
    ctx->phase_row.assign((size_t)P + 1, 0);
    ctx->phase_off.assign((size_t)P + 1, 0);
    CK(cudaMemcpyAsync(ctx->seg_begin.data(), d_seg, 8 * (size_t)(P + 2), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(ctx->phase_row.data(), d_prow, 8 * (size_t)(P + 1), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(ctx->phase_off.data(), d_poff, 8 * (size_t)(P + 1), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaGetLastError());
    ctx->seg_begin[(size_t)P + 1] = n;
    const int neas = eas_neas(ctx->tech);
    const int hist_w = neas ? hex8eas::hist_size(neas) : mulf_points(ctx) * MULF_HIST;
    int rc = permute_array(ctx, ctx->dconn, n, ctx->nn, perm);
    if (!rc)
    {
      if (ctx->d_ele_id)
        rc = permute_array(ctx, ctx->d_ele_id, n, 1, perm);
      else
      {
        CK(cudaMalloc((void**)&ctx->d_ele_id, 4 * (size_t)n));
        CK(cudaMemcpyAsync(ctx->d_ele_id, perm, 4 * (size_t)n, cudaMemcpyDeviceToDevice, st));
      }
    }
    if (!rc) rc = permute_array(ctx, ctx->d_per_ele, n, 1, perm);
    if (!rc) rc = permute_array(ctx, ctx->d_fibers, n, 3, perm);
    if (!rc) rc = permute_array(ctx, ctx->d_gp_hist, n, mulf_points(ctx) * MULF_HIST, perm);
    if (!rc) rc = permute_array(ctx, ctx->d_eas_hist, n, neas ? hex8eas::hist_size(neas) : 0, perm);
    if (!rc) rc = permute_array(ctx, ctx->d_eas_t0, n, hex8eas::T0_STRIDE, perm);
    if (!rc) rc = permute_array(ctx, ctx->d_hist_old, n, hist_w, perm);
    if (!rc) rc = permute_array(ctx, ctx->d_mat_hist, n, ctx->ngp * mat_hist_per_gp(ctx), perm);
    CK(cudaStreamSynchronize(st));
    cudaFree(d_prow), cudaFree(d_poff), cudaFree(d_seg), cudaFree(d_spread), cudaFree(k0), cudaFree(k1);
    cudaFree(i0), cudaFree(perm), cudaFree(tmp);
    if (rc) return B200S_ERR_CUDA;
    ctx->segmented = true;
    return 0;
  }

  // element -> CSR slot map of this context's element block against the graph in ctx->d_colind / d_node_row
  static int install_block_map(b200s_ctx* ctx)
  {
    if (order_elements_for_graph(ctx)) return B200S_ERR_CUDA;
    ctx->pos_stride = ctx->celltype == B200S_HEX27 ? hex27::POS_STRIDE : ctx->nn * ctx->nn;
    if (to_device(ctx, &ctx->d_pos, (const int32_t*)nullptr, ctx->n_col_eles * ctx->pos_stride))
      return B200S_ERR_CUDA;
    if (!ctx->d_nm_flag && to_device(ctx, &ctx->d_nm_flag, (const int*)nullptr, 1)) return B200S_ERR_CUDA;
    CK(cudaMemsetAsync(ctx->d_nm_flag, 0, sizeof(int), ctx->stream));
    k_build_scatter_map<<<nblk(ctx->n_col_eles * ctx->nn), 256, 0, ctx->stream>>>(ctx->nn,
        ctx->pos_stride, ctx->n_col_eles, ctx->dconn, ctx->d_colind, ctx->d_node_row, ctx->d_node_colbase, ctx->d_pos,
        ctx->d_nm_flag, ctx->nn);
    if (ctx->celltype == B200S_HEX20 || ctx->celltype == B200S_HEX18)
    {
      // hex20 / hex18 blocks run through the hex27 kernel: connectivity padded to 27 nodes (padding = the element's node 0,
      // whose coordinates meet zero shape derivatives) and the 27 x 27 scatter map with -1 for the padding rows / columns
      if (to_device(ctx, &ctx->d_conn27, (const int32_t*)nullptr, ctx->n_col_eles * 27) ||
          to_device(ctx, &ctx->d_pos27, (const int32_t*)nullptr, ctx->n_col_eles * hex27::POS_STRIDE))
        return B200S_ERR_CUDA;
      k_pad_conn<<<nblk(ctx->n_col_eles * 27), 256, 0, ctx->stream>>>(ctx->n_col_eles, ctx->nn, 27, ctx->dconn, ctx->d_conn27);
      k_build_scatter_map<<<nblk(ctx->n_col_eles * 27), 256, 0, ctx->stream>>>(27, hex27::POS_STRIDE, ctx->n_col_eles,
          ctx->d_conn27, ctx->d_colind, ctx->d_node_row, ctx->d_node_colbase, ctx->d_pos27, ctx->d_nm_flag, ctx->nn);
      ctx->launches += 2;
    }
    ctx->launches += 1;
    int flag = 0;
    CK(cudaMemcpyAsync(&flag, ctx->d_nm_flag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaGetLastError());
    if (flag)
    {
      dfree(ctx->d_pos);
      return fail(ctx, B200S_ERR_UNSUPPORTED,
          "graph layout outside the fast assembly path (flag 8: an element column is missing / not contiguous in "
          "an owned row)");
    }
    return B200S_OK;
  }

  static int install_graph(b200s_ctx* ctx, int64_t n_rows, const int64_t* d_rowptr,
      const int32_t* d_colind, const int32_t* d_dof_row_lid, const int32_t* d_dof_col_lid)
  {
    if (ctx->primary) return fail(ctx, B200S_ERR_INVALID, "an attached element block uses its primary's graph");
    delete ctx->sym;  // transfer tables of the previous graph
    ctx->sym = nullptr;
    const int64_t ncn = ctx->n_col_nodes;
    int64_t nnz = 0;
    CK(cudaMemcpyAsync(&nnz, d_rowptr + n_rows, sizeof(int64_t), cudaMemcpyDefault, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->n_rows = n_rows, ctx->nnz = nnz;
    if (to_device(ctx, &ctx->d_node_row, (const NodeRow*)nullptr, ncn)) return B200S_ERR_CUDA;
    if (to_device(ctx, &ctx->d_node_colbase, (const int32_t*)nullptr, ncn)) return B200S_ERR_CUDA;
    if (ctx->zero_stream) cudaStreamSynchronize(ctx->zero_stream);
    dfree(ctx->d_vals_buf[1]);
    if (to_device(ctx, &ctx->d_vals_buf[0], (const double*)nullptr, nnz)) return B200S_ERR_CUDA;
    if (ctx->double_buffer && to_device(ctx, &ctx->d_vals_buf[1], (const double*)nullptr, nnz))
      return B200S_ERR_CUDA;
    ctx->vals_cur = 0, ctx->d_vals = ctx->d_vals_buf[0];
    ctx->vals_zeroed[0] = ctx->vals_zeroed[1] = false;
    if (to_device(ctx, &ctx->d_f, (const double*)nullptr, n_rows)) return B200S_ERR_CUDA;
    if (to_device(ctx, &ctx->d_u_row, (const double*)nullptr, n_rows)) return B200S_ERR_CUDA;
    CK(cudaMemsetAsync(ctx->d_u_row, 0, sizeof(double) * n_rows, ctx->stream));
    int* d_flag = nullptr;
    CK(cudaMalloc((void**)&d_flag, sizeof(int)));
    CK(cudaMemsetAsync(d_flag, 0, sizeof(int), ctx->stream));
    k_node_rows<<<nblk(ncn), 256, 0, ctx->stream>>>(ncn, d_dof_row_lid, d_dof_col_lid, d_rowptr,
        n_rows, ctx->d_node_row, ctx->d_node_colbase, d_flag);
    ctx->launches += 1;
    int flag = 0;
    CK(cudaMemcpyAsync(&flag, d_flag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    cudaFree(d_flag);
    CK(cudaGetLastError());
    // buffers of the optional actions are sized by the graph: drop stale ones
    dfree(ctx->d_mass), dfree(ctx->d_f_inertia), dfree(ctx->d_acc_row), dfree(ctx->d_xmat);
    dfree(ctx->d_la_x), dfree(ctx->d_la_y), dfree(ctx->d_toggle);
    dfree(ctx->d_fext), dfree(ctx->d_nm_row), dfree(ctx->d_du_row);
    for (auto& v : ctx->d_cg) dfree(v);
    dfree(ctx->d_cg_hist);
    ctx->cg_hist_len = 0;
    if (flag)
    {
      dfree(ctx->d_pos);
      ctx->nnz = 0;
      return fail(ctx, B200S_ERR_UNSUPPORTED,
          "graph layout outside the fast assembly path (flags 0x%x: 1 col dofs of a node not "
          "contiguous, 2 row dofs not contiguous, 4 unequal row lengths of a node)",
          flag);
    }
    {
      const int rc = install_block_map(ctx);
      if (rc)
      {
        ctx->nnz = 0;
        return rc;
      }
    }
    // matrix column count and the node-blocked pattern test of the fast multiply
    std::vector<int32_t> cb(ncn);
    CK(cudaMemcpy(cb.data(), ctx->d_node_colbase, sizeof(int32_t) * ncn, cudaMemcpyDeviceToHost));
    int32_t mx = -1;
    for (int32_t v : cb) mx = v > mx ? v : mx;
    ctx->n_mat_cols = (int64_t)mx + 3;
    if (to_device(ctx, &ctx->d_row_diagcol, (const int32_t*)nullptr, n_rows)) return B200S_ERR_CUDA;
    CK(cudaMemsetAsync(ctx->d_row_diagcol, 0xff, sizeof(int32_t) * (size_t)n_rows, ctx->stream));
    k_row_diag_col<<<nblk(3 * ncn), 256, 0, ctx->stream>>>(ncn, ctx->d_node_row, ctx->d_node_colbase,
        ctx->d_row_diagcol);
    ctx->launches += 1;
    int* d_f3 = nullptr;
    int f3 = 0;
    CK(cudaMalloc((void**)&d_f3, sizeof(int)));
    CK(cudaMemsetAsync(d_f3, 0, sizeof(int), ctx->stream));
    k_check_node3<<<nblk(32 * ncn), 256, 0, ctx->stream>>>(ncn, ctx->d_node_row, d_colind, d_f3);
    ctx->launches += 1;
    CK(cudaMemcpyAsync(&f3, d_f3, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    cudaFree(d_f3);
    ctx->node3_ok = (f3 == 0);
    dfree(ctx->d_blkcol);
    if (ctx->node3_ok)
    {
      if (to_device(ctx, &ctx->d_blkcol, (const int32_t*)nullptr, nnz / 9 + 1)) return B200S_ERR_CUDA;
      k_build_blkcol<<<nblk(32 * ncn), 256, 0, ctx->stream>>>(ncn, ctx->d_node_row, d_colind, ctx->d_blkcol);
      ctx->launches += 1;
      CK(cudaStreamSynchronize(ctx->stream));
    }
    return B200S_OK;
  }

  int b200s_set_graph(b200s_ctx* ctx, int64_t n_rows, const int64_t* rowptr, const int32_t* colind,
      const int32_t* dof_row_lid, const int32_t* dof_col_lid)
  {
    if (!ctx || !rowptr || n_rows < 0) return B200S_ERR_INVALID;
    if (ctx->celltype < 0) return fail(ctx, B200S_ERR_INVALID, "set_mesh must precede set_graph");
    // empty graphs (no row dofs / no entries) are valid; colind and the dof maps may then be absent
    if ((ctx->n_col_nodes > 0 && (!dof_row_lid || !dof_col_lid)) || (n_rows > 0 && !colind))
      return fail(ctx, B200S_ERR_INVALID, "set_graph: colind / dof_row_lid / dof_col_lid missing");
    if (ctx->primary) return fail(ctx, B200S_ERR_INVALID, "an attached element block uses its primary's graph");
    cudaSetDevice(ctx->device);
    // stage the caller's (host or device) arrays on the device; they are only needed to build the map
    // the graph stays on the device: apply_dirichlet / multiply need it after the map is built
    int32_t *d_r = nullptr, *d_c = nullptr;
    int64_t nnz = 0;
    CK(cudaMemcpy(&nnz, rowptr + n_rows, sizeof(int64_t), cudaMemcpyDefault));
    int rc = 0;
    rc |= to_device(ctx, &ctx->d_rowptr, rowptr, n_rows + 1);
    rc |= to_device(ctx, &ctx->d_colind, colind, nnz);
    rc |= to_device(ctx, &d_r, dof_row_lid, 3 * ctx->n_col_nodes);
    rc |= to_device(ctx, &d_c, dof_col_lid, 3 * ctx->n_col_nodes);
    if (!rc) rc = install_graph(ctx, n_rows, ctx->d_rowptr, ctx->d_colind, d_r, d_c);
    cudaFree(d_r), cudaFree(d_c);
    return rc ? (rc < 0 ? rc : B200S_ERR_CUDA) : B200S_OK;
  }

  int b200s_build_graph(b200s_ctx* ctx, int64_t n_row_nodes, const int32_t* node_row_lid,
      const int32_t* node_col_lid, int64_t* rowptr_out, int32_t* colind_out, int64_t* nnz_out)
  {
    if (!ctx || n_row_nodes < 0) return B200S_ERR_INVALID;
    if (ctx->celltype < 0) return fail(ctx, B200S_ERR_INVALID, "set_mesh must precede build_graph");
    // a processor without row nodes (ghost elements only) or without any node is valid: empty graph
    if (ctx->n_col_nodes > 0 && (!node_row_lid || !node_col_lid))
      return fail(ctx, B200S_ERR_INVALID, "build_graph: node_row_lid / node_col_lid missing");
    if (ctx->primary) return fail(ctx, B200S_ERR_INVALID, "an attached element block uses its primary's graph");
    cudaSetDevice(ctx->device);
    cudaStream_t st = ctx->stream;
    const int nn = ctx->nn;
    const int64_t ncn = ctx->n_col_nodes, n_keys = ctx->n_col_eles * nn * nn;
    if (colind_out == nullptr || ctx->d_rowptr == nullptr)
    {
      int32_t *d_r = nullptr, *d_c = nullptr;
      if (to_device(ctx, &d_r, node_row_lid, ncn) || to_device(ctx, &d_c, node_col_lid, ncn))
        return B200S_ERR_CUDA;
      unsigned long long *k0 = nullptr, *k1 = nullptr;
      int64_t* d_nsel = nullptr;
      CK(cudaMalloc((void**)&k0, sizeof(unsigned long long) * n_keys));
      CK(cudaMalloc((void**)&k1, sizeof(unsigned long long) * n_keys));
      CK(cudaMalloc((void**)&d_nsel, sizeof(int64_t)));
      k_graph_keys<<<nblk(n_keys), 256, 0, st>>>(nn, ctx->n_col_eles, ctx->dconn, d_r, d_c, k0);
      size_t tmp_bytes = 0, tb2 = 0;
      cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, k0, k1, n_keys, 0, 64, st);
      cub::DeviceSelect::Unique(nullptr, tb2, k1, k0, d_nsel, n_keys, st);
      if (tb2 > tmp_bytes) tmp_bytes = tb2;
      void* tmp = nullptr;
      CK(cudaMalloc(&tmp, tmp_bytes));
      cub::DeviceRadixSort::SortKeys(tmp, tmp_bytes, k0, k1, n_keys, 0, 64, st);
      cub::DeviceSelect::Unique(tmp, tmp_bytes, k1, k0, d_nsel, n_keys, st);
      int64_t nsel = 0;
      CK(cudaMemcpyAsync(&nsel, d_nsel, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
      CK(cudaStreamSynchronize(st));
      // per row node neighbour counts -> nodeptr
      int* d_cnt = nullptr;
      int64_t* d_nodeptr = nullptr;
      CK(cudaMalloc((void**)&d_cnt, sizeof(int) * (n_row_nodes + 1)));
      CK(cudaMalloc((void**)&d_nodeptr, sizeof(int64_t) * (n_row_nodes + 1)));
      CK(cudaMemsetAsync(d_cnt, 0, sizeof(int) * (n_row_nodes + 1), st));
      k_graph_count<<<nblk(nsel), 256, 0, st>>>(nsel, k0, d_cnt);
      size_t tb3 = 0;
      cub::DeviceScan::ExclusiveSum(nullptr, tb3, d_cnt, d_nodeptr, n_row_nodes + 1, st);
      void* tmp3 = nullptr;
      CK(cudaMalloc(&tmp3, tb3));
      cub::DeviceScan::ExclusiveSum(tmp3, tb3, d_cnt, d_nodeptr, n_row_nodes + 1, st);
      dfree(ctx->d_rowptr), dfree(ctx->d_colind);
      CK(cudaMalloc((void**)&ctx->d_rowptr, sizeof(int64_t) * (3 * n_row_nodes + 1)));
      k_graph_rowptr<<<nblk(n_row_nodes + 1), 256, 0, st>>>(n_row_nodes, d_nodeptr, ctx->d_rowptr);
      int64_t nnz = 0;
      CK(cudaMemcpyAsync(&nnz, ctx->d_rowptr + 3 * n_row_nodes, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
      CK(cudaStreamSynchronize(st));
      CK(cudaMalloc((void**)&ctx->d_colind, sizeof(int32_t) * (size_t)(nnz > 0 ? nnz : 1)));
      k_graph_colind<<<nblk(nsel), 256, 0, st>>>(nsel, k0, d_nodeptr, ctx->d_colind);
      ctx->launches += 4;
      CK(cudaStreamSynchronize(st));
      CK(cudaGetLastError());
      cudaFree(k0), cudaFree(k1), cudaFree(d_nsel), cudaFree(tmp), cudaFree(tmp3), cudaFree(d_cnt);
      cudaFree(d_nodeptr);
      // install: dof lids are 3*node lid + d
      std::vector<int32_t> hr(3 * ncn), hc(3 * ncn), nr(ncn), nc(ncn);
      CK(cudaMemcpy(nr.data(), d_r, sizeof(int32_t) * ncn, cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(nc.data(), d_c, sizeof(int32_t) * ncn, cudaMemcpyDeviceToHost));
      for (int64_t k = 0; k < ncn; ++k)
        for (int d = 0; d < 3; ++d)
        {
          hr[3 * k + d] = nr[k] < 0 ? -1 : 3 * nr[k] + d;
          hc[3 * k + d] = 3 * nc[k] + d;
        }
      int32_t *d_dr = nullptr, *d_dc = nullptr;
      int rc = to_device(ctx, &d_dr, hr.data(), 3 * ncn) | to_device(ctx, &d_dc, hc.data(), 3 * ncn);
      if (!rc) rc = install_graph(ctx, 3 * n_row_nodes, ctx->d_rowptr, ctx->d_colind, d_dr, d_dc);
      cudaFree(d_dr), cudaFree(d_dc), cudaFree(d_r), cudaFree(d_c);
      if (rc) return rc < 0 ? rc : B200S_ERR_CUDA;
    }
    if (nnz_out) *nnz_out = ctx->nnz;
    if (rowptr_out)
      CK(cudaMemcpy(rowptr_out, ctx->d_rowptr, sizeof(int64_t) * (ctx->n_rows + 1), cudaMemcpyDefault));
    if (colind_out)
      CK(cudaMemcpy(colind_out, ctx->d_colind, sizeof(int32_t) * ctx->nnz, cudaMemcpyDefault));
    return B200S_OK;
  }

  // Element block `ctx` (its own set_mesh / set_element_properties / set_material; the SAME column nodes as the
  // primary) assembles into the primary's system matrix.
  int b200s_attach_block(b200s_ctx* ctx, b200s_ctx* primary)
  {
    if (!ctx || !primary || ctx == primary) return B200S_ERR_INVALID;
    if (primary->primary) return fail(ctx, B200S_ERR_INVALID, "attach to the primary context, not to another attached block");
    if (ctx->celltype < 0) return fail(ctx, B200S_ERR_INVALID, "set_mesh must precede attach_block");
    if (!primary->d_rowptr || !primary->d_node_row || primary->nnz <= 0)
      return fail(ctx, B200S_ERR_INVALID, "the primary context has no graph yet (set_graph / build_graph first)");
    if (ctx->device != primary->device || ctx->n_col_nodes != primary->n_col_nodes)
      return fail(ctx, B200S_ERR_INVALID,
          "an attached block lives on the primary's device and uses the primary's column nodes (%lld here, %lld there)",
          (long long)ctx->n_col_nodes, (long long)primary->n_col_nodes);
    cudaSetDevice(ctx->device);
    CK(cudaStreamSynchronize(ctx->stream));
    if (ctx->primary)
      detach_shared(ctx);
    else
    {
      // drop an own graph, if any
      if (ctx->zero_stream) cudaStreamSynchronize(ctx->zero_stream);
      dfree(ctx->d_rowptr), dfree(ctx->d_colind), dfree(ctx->d_node_row), dfree(ctx->d_node_colbase);
      dfree(ctx->d_row_diagcol), dfree(ctx->d_blkcol), dfree(ctx->d_vals_buf[0]), dfree(ctx->d_vals_buf[1]);
      dfree(ctx->d_f), dfree(ctx->d_u_row), dfree(ctx->d_u_col);
      if (ctx->comm && ctx->nccl.CommDestroy) ctx->nccl.CommDestroy(ctx->comm);
      ctx->comm = nullptr, ctx->n_nbr = 0;
    }
    ctx->primary = primary;
    ctx->stream = primary->stream;
    ctx->d_rowptr = primary->d_rowptr, ctx->d_colind = primary->d_colind;
    ctx->d_node_row = primary->d_node_row, ctx->d_node_colbase = primary->d_node_colbase;
    ctx->d_row_diagcol = primary->d_row_diagcol, ctx->d_blkcol = primary->d_blkcol;
    ctx->d_vals = primary->d_vals, ctx->d_f = primary->d_f;
    ctx->d_u_row = primary->d_u_row, ctx->d_u_col = primary->d_u_col;
    ctx->n_rows = primary->n_rows, ctx->nnz = primary->nnz;
    ctx->n_mat_cols = primary->n_mat_cols, ctx->node3_ok = primary->node3_ok;
    ctx->rank = primary->rank, ctx->nranks = primary->nranks;
    const int rc = install_block_map(ctx);
    if (rc)
    {
      detach_shared(ctx);
      return rc;
    }
    return B200S_OK;
  }

  // ---- halo ----------------------------------------------------------------------------------------
  static int load_nccl(b200s_ctx* ctx)
  {
    NcclApi& n = ctx->nccl;
    if (n.handle) return 0;
    n.handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!n.handle) return fail(ctx, B200S_ERR_COMM, "dlopen(libnccl.so.2): %s", dlerror());
#define SYM(field, name)                                             \
  *(void**)(&n.field) = dlsym(n.handle, name);                       \
  if (!n.field) return fail(ctx, B200S_ERR_COMM, "missing %s", name)
    SYM(GetUniqueId, "ncclGetUniqueId");
    SYM(CommInitRank, "ncclCommInitRank");
    SYM(CommDestroy, "ncclCommDestroy");
    SYM(Send, "ncclSend");
    SYM(Recv, "ncclRecv");
    SYM(GroupStart, "ncclGroupStart");
    SYM(GroupEnd, "ncclGroupEnd");
    SYM(AllReduce, "ncclAllReduce");
    SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
    return 0;
  }

  int b200s_nccl_unique_id(b200s_ctx* ctx, char* out128)
  {
    if (!ctx || !out128) return B200S_ERR_INVALID;
    if (load_nccl(ctx)) return B200S_ERR_COMM;
    Uid id;
    const int rc = ctx->nccl.GetUniqueId(&id);
    if (rc) return fail(ctx, B200S_ERR_COMM, "ncclGetUniqueId: %s", ctx->nccl.GetErrorString(rc));
    memcpy(out128, id.internal, 128);
    return B200S_OK;
  }

  int b200s_comm_init_nccl(b200s_ctx* ctx, const char* id128)
  {
    if (!ctx || !id128) return B200S_ERR_INVALID;
    if (load_nccl(ctx)) return B200S_ERR_COMM;
    cudaSetDevice(ctx->device);
    Uid id;
    memcpy(id.internal, id128, 128);
    const int rc = ctx->nccl.CommInitRank(&ctx->comm, ctx->nranks, id, ctx->rank);
    if (rc) return fail(ctx, B200S_ERR_COMM, "ncclCommInitRank: %s", ctx->nccl.GetErrorString(rc));
    return B200S_OK;
  }

  int b200s_set_halo(b200s_ctx* ctx, int n_nbr, const int32_t* peer, const int64_t* send_offsets,
      const int32_t* send_row_lids, const int64_t* recv_offsets, const int32_t* recv_col_lids)
  {
    if (!ctx || n_nbr < 0) return B200S_ERR_INVALID;
    cudaSetDevice(ctx->device);
    ctx->n_nbr = n_nbr;
    ctx->peer.assign(peer, peer + n_nbr);
    ctx->send_off.assign(send_offsets, send_offsets + n_nbr + 1);
    ctx->recv_off.assign(recv_offsets, recv_offsets + n_nbr + 1);
    const int64_t ns = ctx->send_off[n_nbr], nr = ctx->recv_off[n_nbr];
    if (to_device(ctx, &ctx->d_send_lid, send_row_lids, ns)) return B200S_ERR_CUDA;
    if (to_device(ctx, &ctx->d_recv_lid, recv_col_lids, nr)) return B200S_ERR_CUDA;
    if (to_device(ctx, &ctx->d_sendbuf, (const double*)nullptr, 3 * ns)) return B200S_ERR_CUDA;
    if (to_device(ctx, &ctx->d_recvbuf, (const double*)nullptr, 3 * nr)) return B200S_ERR_CUDA;
    CK(cudaStreamSynchronize(ctx->stream));
    return B200S_OK;
  }

  static inline bool action_has_stiffness(int a)
  {
    return a == B200S_ACTION_NLNSTIFF || a == B200S_ACTION_NLNSTIFFMASS || a == B200S_ACTION_NLNSTIFFLMASS;
  }
  static inline bool action_has_mass(int a)
  {
    return a == B200S_ACTION_NLNSTIFFMASS || a == B200S_ACTION_NLNSTIFFLMASS ||
           a == B200S_ACTION_INTERNALINERTIAFORCE;
  }
  static int import_row_to_col(b200s_ctx* ctx, const double* d_src_row, double* d_dst, bool to_matcol);
  static double* matrix_values(b200s_ctx* ctx, int which);

  // mass matrix (second system matrix, same graph) and inertia force M_e * acc of the mass actions
  // (4C_solid_ele_evaluate.cpp:116-159).  The "acceleration" state is optional: without it only the
  // mass matrix is produced (the reference skips the product for an empty elevec2, :36-47).
  static int launch_mass(b200s_ctx* ctx, int action)
  {
    cudaStream_t st = ctx->stream;
    const bool assemble = action != B200S_ACTION_INTERNALINERTIAFORCE;
    if (assemble)
    {
      if (!ctx->d_mass && to_device(ctx, &ctx->d_mass, (const double*)nullptr, ctx->nnz)) return B200S_ERR_CUDA;
      CK(cudaMemsetAsync(ctx->d_mass, 0, sizeof(double) * (size_t)ctx->nnz, st));
    }
    if (!ctx->d_f_inertia && to_device(ctx, &ctx->d_f_inertia, (const double*)nullptr, ctx->n_rows))
      return B200S_ERR_CUDA;
    CK(cudaMemsetAsync(ctx->d_f_inertia, 0, sizeof(double) * (size_t)ctx->n_rows, st));
    if (ctx->d_acc_row)
    {
      if (!ctx->d_acc_col && to_device(ctx, &ctx->d_acc_col, (const double*)nullptr, 3 * ctx->n_col_nodes))
        return B200S_ERR_CUDA;
      const int rc = import_row_to_col(ctx, ctx->d_acc_row, ctx->d_acc_col, false);
      if (rc) return rc;
    }
    MassArgs M;
    M.X = ctx->dX, M.conn = ctx->dconn, M.pos = ctx->d_pos, M.node_row = ctx->d_node_row;
    M.acc = ctx->d_acc_row ? ctx->d_acc_col : nullptr;
    M.N_mass = ctx->d_N_mass, M.dN_mass = ctx->d_dN_mass, M.w_mass = ctx->d_w_mass;
    M.dN_gp = ctx->d_dN_gp, M.w = ctx->d_w;
    M.density = ctx->density, M.n_ele = ctx->n_col_eles;
    M.nn = ctx->nn, M.ngp = ctx->ngp, M.ngp_mass = ctx->ngp_mass, M.pos_stride = ctx->pos_stride;
    // internalinertiaforce uses the consistent element mass matrix (:140-158)
    M.lumped = action == B200S_ACTION_NLNSTIFFLMASS;
    M.mass_vals = assemble ? ctx->d_mass : nullptr;
    M.f_inertia = ctx->d_f_inertia;
    const unsigned grid = nblk(ctx->n_col_eles * ctx->nn, 128);
    if (grid)
    {
      switch (ctx->celltype)
      {
        case B200S_HEX8:
          k_mass_assemble<8><<<grid, 128, 0, st>>>(M);
          break;
        case B200S_HEX27:
          k_mass_assemble<27><<<grid, 128, 0, st>>>(M);
          break;
        case B200S_TET10:
          k_mass_assemble<10><<<grid, 128, 0, st>>>(M);
          break;
        case B200S_TET4:
          k_mass_assemble<4><<<grid, 128, 0, st>>>(M);
          break;
        case B200S_HEX20:
          k_mass_assemble<20><<<grid, 128, 0, st>>>(M);
          break;
        case B200S_WEDGE6:
          k_mass_assemble<6><<<grid, 128, 0, st>>>(M);
          break;
        case B200S_PYRAMID5:
          k_mass_assemble<5><<<grid, 128, 0, st>>>(M);
          break;
        case B200S_HEX18:
          k_mass_assemble<18><<<grid, 128, 0, st>>>(M);
          break;
      }
      ctx->launches += 1;
    }
    return B200S_OK;
  }

  // Import of a row-map vector to the column map (Epetra_Import of Discretization::set_state,
  // 4C_fem_discretization.cpp:612-665; the same import Epetra_CrsMatrix::Multiply does for x).
  // to_matcol == false: dst is indexed by column-node lid (3 per node);
  // to_matcol == true : dst is indexed by the matrix-local column id (node_colbase + d).
  // Two halves writing disjoint entries of dst: the ghost part (pack, ncclSend/Recv group, unpack) and the owned
  // part; evaluate runs the first on a side stream while the elements without ghost nodes are evaluated.
  static int import_ghosts(b200s_ctx* ctx, const double* d_src_row, double* d_dst, bool to_matcol, cudaStream_t st)
  {
    if (ctx->n_nbr <= 0) return B200S_OK;
    if (!ctx->comm) return fail(ctx, B200S_ERR_COMM, "halo set but no communicator");
    const int64_t ns = ctx->send_off[ctx->n_nbr], nr = ctx->recv_off[ctx->n_nbr];
    if (ns) k_pack_halo<<<nblk(3 * ns), 256, 0, st>>>(ns, ctx->d_send_lid, d_src_row, ctx->d_sendbuf);
    ctx->nccl.GroupStart();
    for (int n = 0; n < ctx->n_nbr; ++n)
    {
      const int64_t s0 = ctx->send_off[n], s1 = ctx->send_off[n + 1];
      const int64_t r0 = ctx->recv_off[n], r1 = ctx->recv_off[n + 1];
      if (s1 > s0)
        ctx->nccl.Send(ctx->d_sendbuf + 3 * s0, (size_t)(3 * (s1 - s0)), /*ncclDouble*/ 8,
            ctx->peer[n], ctx->comm, st);
      if (r1 > r0)
        ctx->nccl.Recv(ctx->d_recvbuf + 3 * r0, (size_t)(3 * (r1 - r0)), 8, ctx->peer[n],
            ctx->comm, st);
    }
    const int rc = ctx->nccl.GroupEnd();
    if (rc) return fail(ctx, B200S_ERR_COMM, "halo exchange: %s", ctx->nccl.GetErrorString(rc));
    if (nr)
    {
      if (to_matcol)
        k_unpack_halo_matcol<<<nblk(3 * nr), 256, 0, st>>>(
            nr, ctx->d_recv_lid, ctx->d_node_colbase, ctx->d_recvbuf, d_dst);
      else
        k_unpack_halo<<<nblk(3 * nr), 256, 0, st>>>(nr, ctx->d_recv_lid, ctx->d_recvbuf, d_dst);
    }
    ctx->launches += (ns > 0) + (nr > 0);
    return B200S_OK;
  }
  static int import_owned(b200s_ctx* ctx, const double* d_src_row, double* d_dst, bool to_matcol, cudaStream_t st)
  {
    if (to_matcol)
      k_row_to_matcol<<<nblk(3 * ctx->n_col_nodes), 256, 0, st>>>(
          ctx->n_col_nodes, ctx->d_node_row, ctx->d_node_colbase, d_src_row, d_dst);
    else
      k_row_to_col<<<nblk(3 * ctx->n_col_nodes), 256, 0, st>>>(
          ctx->n_col_nodes, ctx->d_node_row, d_src_row, d_dst, io_rows(ctx));
    ctx->launches += 1;
    return B200S_OK;
  }
  static int import_row_to_col(b200s_ctx* ctx, const double* d_src_row, double* d_dst, bool to_matcol)
  {
    const int rc = import_ghosts(ctx, d_src_row, d_dst, to_matcol, ctx->stream);
    return rc ? rc : import_owned(ctx, d_src_row, d_dst, to_matcol, ctx->stream);
  }

  // ---- evaluate -----------------------------------------------------------------------------------
  // arguments of one element-kernel launch over the device elements [e_begin, e_begin + count)
  static EvalArgs make_eval_args(b200s_ctx* ctx, int64_t e_begin, int64_t count)
  {
    EvalArgs A;
    A.X = ctx->dX, A.u = ctx->d_u_col, A.node_row = ctx->d_node_row;
    A.conn = ctx->dconn + e_begin * ctx->nn;
    A.pos = ctx->d_pos + e_begin * ctx->pos_stride;
    A.per_ele_param = ctx->d_per_ele ? ctx->d_per_ele + e_begin : nullptr;
    A.ele_id = ctx->d_ele_id ? ctx->d_ele_id + e_begin : nullptr;
    A.fibers = ctx->d_fibers ? ctx->d_fibers + 3 * e_begin : nullptr;
    A.gp_hist = ctx->d_gp_hist ? ctx->d_gp_hist + e_begin * mulf_points(ctx) * MULF_HIST : nullptr;
    A.mat_hist = ctx->d_mat_hist ? ctx->d_mat_hist + e_begin * ctx->ngp * mat_hist_per_gp(ctx) : nullptr;
    A.linear = ctx->kinem == B200S_KINEM_LINEAR;
    A.vals = ctx->d_vals, A.f = ctx->d_f, A.err = ctx->d_err, A.mp = ctx->mp;
    A.n_ele = count;
    A.scatter_mode = ctx->scatter_mode;
    A.e_begin = e_begin;
    return A;
  }

  // one element-kernel launch over a range of device elements, bracketed by an event pair of the timing ring
  static int launch_segment(b200s_ctx* ctx, bool stiff, int64_t e_begin, int64_t e_end, bool& first)
  {
    if (e_end <= e_begin) return B200S_OK;
    cudaStream_t st = ctx->stream;
    if (ctx->kev_pending == b200s_ctx::RING) harvest_kernel_times(ctx);
    const EvalArgs A = make_eval_args(ctx, e_begin, e_end - e_begin);
    CK(cudaEventRecord(ctx->kev[2 * ctx->kev_pending], st));
    const int rc = stiff ? launch_eval<true>(ctx, A) : launch_eval<false>(ctx, A);
    if (rc) return rc;
    CK(cudaEventRecord(ctx->kev[2 * ctx->kev_pending + 1], st));
    ctx->kev_first[ctx->kev_pending] = first;
    first = false;
    ctx->kev_pending += 1;
    ctx->last_kernel_launches += 1;
    return B200S_OK;
  }


  // ---- row-element mode: export of the off-rank rows ------------------------------------------------------------------
  __global__ void k_add_exported(int64_t n, const int64_t* pos, const double* buf, double* vals)
  {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) atomicAdd(vals + pos[t], buf[t]);  // two peers may add into one entry
  }
  __global__ void k_add_exported_f(int64_t n, const int32_t* pos, const double* buf, double* f)
  {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) atomicAdd(f + pos[t], buf[t]);
  }
  // the received export rows of peer k (already in the receive buffers) are added to the rank's own rows
  static void add_exported(b200s_ctx* ctx, int k, bool stiff)
  {
    cudaStream_t st = ctx->stream;
    const int64_t v0 = ctx->exp_recv_off[(size_t)k], nv = ctx->exp_recv_off[(size_t)k + 1] - v0;
    const int64_t f0 = ctx->exp_recv_foff[(size_t)k], nf = ctx->exp_recv_foff[(size_t)k + 1] - f0;
    if (stiff && nv > 0) k_add_exported<<<nblk(nv), 256, 0, st>>>(nv, ctx->d_exp_pos + v0, ctx->d_exp_buf + v0, ctx->d_vals);
    if (nf > 0) k_add_exported_f<<<nblk(nf), 256, 0, st>>>(nf, ctx->d_exp_fpos + f0, ctx->d_exp_fbuf + f0, ctx->d_f);
    ctx->launches += (stiff && nv > 0) + (nf > 0);
  }
  static int export_rows(b200s_ctx* ctx, bool stiff)
  {
    cudaStream_t st = ctx->stream;
    const int np = (int)ctx->exp_peer.size();
    ctx->nccl.GroupStart();
    for (int k = 0; k < np; ++k)
    {
      const int64_t sv = ctx->exp_send_val[2 * (size_t)k], nv = ctx->exp_send_val[2 * (size_t)k + 1];
      const int64_t sf = ctx->exp_send_f[2 * (size_t)k], nf = ctx->exp_send_f[2 * (size_t)k + 1];
      const int64_t r0 = ctx->exp_recv_off[(size_t)k], rn = ctx->exp_recv_off[(size_t)k + 1] - r0;
      const int64_t g0 = ctx->exp_recv_foff[(size_t)k], gn = ctx->exp_recv_foff[(size_t)k + 1] - g0;
      // the export rows of one owner are contiguous in the extended values / residual: sent in place
      if (stiff && nv > 0) ctx->nccl.Send(ctx->d_vals + sv, (size_t)nv, /*ncclDouble*/ 8, ctx->exp_peer[(size_t)k], ctx->comm, st);
      if (nf > 0) ctx->nccl.Send(ctx->d_f + sf, (size_t)nf, 8, ctx->exp_peer[(size_t)k], ctx->comm, st);
      if (stiff && rn > 0) ctx->nccl.Recv(ctx->d_exp_buf + r0, (size_t)rn, 8, ctx->exp_peer[(size_t)k], ctx->comm, st);
      if (gn > 0) ctx->nccl.Recv(ctx->d_exp_fbuf + g0, (size_t)gn, 8, ctx->exp_peer[(size_t)k], ctx->comm, st);
    }
    const int rc = ctx->nccl.GroupEnd();
    if (rc) return fail(ctx, B200S_ERR_COMM, "row export: %s", ctx->nccl.GetErrorString(rc));
    for (int k = 0; k < np; ++k) add_exported(ctx, k, stiff);
    return B200S_OK;
  }

  static bool is_host_pointer(const void* p)
  {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess)
    {
      cudaGetLastError();
      return true;
    }
    return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeUnregistered;
  }

  // K_dst / f_dst: optional destinations of the result (b200s_evaluate); a HOST K_dst is filled phase by phase
  // while the later element segments are still being evaluated.
  static int evaluate_impl(b200s_ctx* ctx, int action, double* K_dst, double* f_dst)
  {
    if (!ctx->d_pos || ctx->matkind < 0)
      return fail(ctx, B200S_ERR_INVALID, "set_mesh, set_material and set_graph must precede evaluate");
    cudaStream_t st = ctx->stream;
    if (action < B200S_ACTION_INTERNALFORCE || action > B200S_ACTION_INTERNALINERTIAFORCE)
      return fail(ctx, B200S_ERR_INVALID, "unknown action %d", action);
    const bool stiff = action_has_stiffness(action);
    const bool mass = action_has_mass(action);
    if (mass && !(ctx->density > 0.0))
      return fail(ctx, B200S_ERR_INVALID, "mass actions need b200s_set_density (density > 0)");
    if (material_needs_fibers(ctx) && !ctx->d_fibers)
      return fail(ctx, B200S_ERR_INVALID, "the material has a fibre summand: b200s_set_element_fibers first");
    if (ctx->prestress == B200S_PRESTRESS_MULF && !ctx->d_gp_hist)
      return fail(ctx, B200S_ERR_INVALID, "MULF history missing (set_mesh after set_element_properties?)");
    if (ctx->matkind == B200S_MAT_DRUCKER_PRAGER)
    {
      if (ctx->kinem != B200S_KINEM_LINEAR || ctx->tech != B200S_TECH_NONE || ctx->prestress != B200S_PRESTRESS_NONE)
        return fail(ctx, B200S_ERR_UNSUPPORTED,
            "MAT_Struct_DruckerPrager is a small-strain material: KINEM linear without element technology / prestress");
      if (!ctx->d_mat_hist) return fail(ctx, B200S_ERR_INVALID, "material history missing (set_mesh first)");
    }
    if (ctx->matkind == B200S_MAT_VISCONEOHOOKE || ctx->matkind == B200S_MAT_PLASTIC_NLNLOG)
    {
      const bool eas = ctx->tech == B200S_TECH_EAS_MILD || ctx->tech == B200S_TECH_EAS_FULL;
      if (ctx->kinem != B200S_KINEM_NONLINEAR_TOTLAG || (ctx->tech != B200S_TECH_NONE && !eas) ||
          ctx->prestress != B200S_PRESTRESS_NONE)
        return fail(ctx, B200S_ERR_UNSUPPORTED,
            "finite-strain history materials on the device: KINEM nonlinear, TECH none or eas_mild / eas_full, no prestress");
      if (!ctx->dt_set)
        return fail(ctx, B200S_ERR_INVALID,
            "Time step size not given in evaluation context. (b200s_set_time_step_size)");
      if (!ctx->d_mat_hist) return fail(ctx, B200S_ERR_INVALID, "material history missing (set_mesh first)");
    }
    // an attached block adds into its primary's matrix: the primary's evaluate of this iteration has imported
    // the state and zeroed K / f
    const bool accumulate = ctx->primary != nullptr;
    if (accumulate)
    {
      b200s_ctx* pr = ctx->primary;
      ctx->d_vals = pr->d_vals;  // follows the primary's buffer switch (option double_buffer)
      if (mass) return fail(ctx, B200S_ERR_UNSUPPORTED, "mass actions of an attached block: evaluate them on the primary");
    }
    ctx->last_action = action;
    ctx->last_kernel_launches = 0, ctx->last_pipeline_chunks = 0, ctx->last_halo_overlapped = 0;
    const int64_t n_ele = ctx->n_col_eles;
    const int64_t n_ghost_eles = ctx->segmented ? ctx->seg_begin[1] : 0;
    // pipeline: host destination for K, at least two non-empty phases
    const bool pipeline = stiff && K_dst && ctx->segmented && ctx->n_phases > 1 && ctx->pipeline_phases > 1 &&
                          !mass && !ctx->check_finite && ctx->exp_peer.empty() && is_host_pointer(K_dst);
    // block-upper transfer: symmetric K_e (every device formulation but F-bar), tables usable for this graph
    bool symmetric = pipeline && ctx->symmetric_transfer && !accumulate && ctx->tech != B200S_TECH_FBAR;
    if (symmetric)
    {
      if (!ctx->sym && sym_prepare(ctx)) return fail(ctx, B200S_ERR_CUDA, "symmetric transfer: set-up failed");
      symmetric = ctx->sym->usable;
    }
    // overlap of the ghost import with the elements that do not read it
    const bool overlap = !accumulate && !pipeline && ctx->n_nbr > 0 && ctx->segmented && ctx->overlap_halo &&
                         n_ghost_eles < n_ele;
    nvtxRangePushA("b200s_evaluate");
    CK(cudaEventRecord(ctx->ev[1], st));
    if (!accumulate)
    {
      // set_state("displacement"): ghost import, then owned part
      nvtxRangePushA("halo");
      int rc;
      if (overlap)
      {
        CK(cudaEventRecord(ctx->ev_fork, st));
        CK(cudaStreamWaitEvent(ctx->halo_stream, ctx->ev_fork, 0));
        rc = import_ghosts(ctx, ctx->d_u_row, ctx->d_u_col, false, ctx->halo_stream);
        CK(cudaEventRecord(ctx->ev_halo_done, ctx->halo_stream));
        if (!rc) rc = import_owned(ctx, ctx->d_u_row, ctx->d_u_col, false, st);
        ctx->last_halo_overlapped = 1;
      }
      else
        rc = import_row_to_col(ctx, ctx->d_u_row, ctx->d_u_col, false);
      nvtxRangePop();
      if (rc)
      {
        nvtxRangePop();
        return rc;
      }
    }
    CK(cudaEventRecord(ctx->ev[2], st));
    // stiff.zero(); fint.put_scalar(0)
    if (!accumulate)
    {
      nvtxRangePushA("zero");
      if (stiff)
      {
        const size_t vbytes = sizeof(double) * (size_t)ctx->nnz;
        if (ctx->d_vals_buf[1])
        {
          // this evaluate fills the buffer that was cleared during the previous one (or clears it now,
          // the first time); meanwhile the side stream clears the buffer holding the previous result,
          // once everything enqueued so far (e.g. its copy to the host) has finished
          const int use = ctx->vals_cur ^ 1, idle = ctx->vals_cur;
          CK(cudaEventRecord(ctx->ev_eval_start, st));
          if (ctx->vals_zeroed[use])
            CK(cudaStreamWaitEvent(st, ctx->ev_zero_done[use], 0));
          else
            CK(cudaMemsetAsync(ctx->d_vals_buf[use], 0, vbytes, st));
          CK(cudaStreamWaitEvent(ctx->zero_stream, ctx->ev_eval_start, 0));
          if (ctx->double_buffer == 2)
          {
            // zeros by the copy engine: device-to-device copies from a small (L2 resident) zero block, so that the
            // clearing neither occupies SMs nor competes for their L1TEX with the element kernel
            if (!ctx->d_zero)
            {
              CK(cudaMalloc((void**)&ctx->d_zero, b200s_ctx::ZERO_BLOCK));
              CK(cudaMemsetAsync(ctx->d_zero, 0, b200s_ctx::ZERO_BLOCK, ctx->zero_stream));
            }
            for (size_t o = 0; o < vbytes; o += b200s_ctx::ZERO_BLOCK)
            {
              const size_t nb = vbytes - o < b200s_ctx::ZERO_BLOCK ? vbytes - o : b200s_ctx::ZERO_BLOCK;
              CK(cudaMemcpyAsync((char*)ctx->d_vals_buf[idle] + o, ctx->d_zero, nb, cudaMemcpyDeviceToDevice,
                  ctx->zero_stream));
            }
          }
          else
            CK(cudaMemsetAsync(ctx->d_vals_buf[idle], 0, vbytes, ctx->zero_stream));
          CK(cudaEventRecord(ctx->ev_zero_done[idle], ctx->zero_stream));
          ctx->vals_zeroed[idle] = true, ctx->vals_zeroed[use] = false;
          ctx->vals_cur = use, ctx->d_vals = ctx->d_vals_buf[use];
        }
        else
          CK(cudaMemsetAsync(ctx->d_vals, 0, vbytes, st));
      }
      CK(cudaMemsetAsync(ctx->d_f, 0, sizeof(double) * (size_t)ctx->n_rows, st));
      nvtxRangePop();
    }
    // reset word from a constant pinned source: h_err itself may still be the target of the previous
    // (asynchronous) evaluate's read-back
    CK(cudaMemcpyAsync(ctx->d_err, ctx->h_err_init, sizeof(ErrWord), cudaMemcpyHostToDevice, st));
    CK(cudaEventRecord(ctx->ev[3], st));
    nvtxRangePushA("element kernels");
    bool first = true;
    int rc = B200S_OK;
    if (overlap)
    {
      // elements without ghost nodes while the halo travels, the others once it has arrived
      rc = launch_segment(ctx, stiff, n_ghost_eles, n_ele, first);
      CK(cudaStreamWaitEvent(st, ctx->ev_halo_done, 0));
      if (!rc) rc = launch_segment(ctx, stiff, 0, n_ghost_eles, first);
    }
    else if (pipeline)
    {
      // segment 0 (ghost-node elements) first, then phase by phase; the rows of phase k are complete after
      // segment k + 1 and leave on the copy stream while the next segments are evaluated
      rc = launch_segment(ctx, stiff, 0, n_ghost_eles, first);
      for (int k = 0; k < ctx->n_phases && !rc; ++k)
      {
        rc = launch_segment(ctx, stiff, ctx->seg_begin[(size_t)k + 1], ctx->seg_begin[(size_t)k + 2], first);
        const int64_t o0 = ctx->phase_off[(size_t)k], o1 = ctx->phase_off[(size_t)k + 1];
        if (rc || o1 <= o0) continue;
        if (symmetric)
        {
          // only the blocks with column node >= row node travel; the host threads write the full rows (below)
          SymTransfer& T = *ctx->sym;
          const int64_t n0 = T.phase_node[(size_t)k], n1 = T.phase_node[(size_t)k + 1];
          const int64_t p0 = T.up_start[(size_t)n0], p1 = T.up_start[(size_t)n1];
          k_pack_upper<<<nblk(32 * (n1 - n0)), 256, 0, st>>>(n0, n1, T.d_csr_start, T.d_nb, T.d_dp, T.d_up_start,
              ctx->d_vals, T.d_packed);
          ctx->launches += 1;
          CK(cudaEventRecord(ctx->ev_phase[k], st));
          CK(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_phase[k], 0));
          CK(cudaMemcpyAsync(T.h_packed + p0, T.d_packed + p0, sizeof(double) * (size_t)(p1 - p0), cudaMemcpyDeviceToHost,
              ctx->copy_stream));
          CK(cudaEventRecord(T.ev_arrive[k], ctx->copy_stream));
          ctx->last_pipeline_chunks += 1;
          continue;
        }
        CK(cudaEventRecord(ctx->ev_phase[k], st));
        CK(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_phase[k], 0));
        CK(cudaMemcpyAsync(K_dst + o0, ctx->d_vals + o0, sizeof(double) * (size_t)(o1 - o0), cudaMemcpyDeviceToHost,
            ctx->copy_stream));
        ctx->last_pipeline_chunks += 1;
      }
      CK(cudaEventRecord(ctx->ev_copy_done, ctx->copy_stream));
    }
    else
      rc = launch_segment(ctx, stiff, 0, n_ele, first);
    nvtxRangePop();
    if (rc)
    {
      nvtxRangePop();
      return rc;
    }
    if (!ctx->exp_peer.empty() && ctx->comm)
    {
      // complete() of the FE matrix: export rows -> owners, Add mode (4C_linalg_sparsematrix.cpp:598-630)
      nvtxRangePushA("row export");
      const int erc = export_rows(ctx, stiff);
      nvtxRangePop();
      if (erc)
      {
        nvtxRangePop();
        return erc;
      }
    }
    if (mass)
    {
      const int mrc = launch_mass(ctx, action);
      if (mrc)
      {
        nvtxRangePop();
        return mrc;
      }
    }
    if (ctx->check_finite)
    {
      if (stiff)
        k_count_nonfinite<<<ctx->n_sm * 8, 256, 0, st>>>(ctx->nnz, ctx->d_vals, &ctx->d_err->n_nonfinite);
      k_count_nonfinite<<<ctx->n_sm, 256, 0, st>>>(ctx->n_rows, ctx->d_f, &ctx->d_err->n_nonfinite);
      ctx->launches += 1 + stiff;
    }
    CK(cudaEventRecord(ctx->ev[4], st));
    CK(cudaMemcpyAsync(ctx->h_err, ctx->d_err, sizeof(ErrWord), cudaMemcpyDeviceToHost, st));
    // results to the caller's buffers
    if (K_dst && stiff && K_dst != ctx->d_vals)
    {
      nvtxRangePushA("d2h");
      if (pipeline)
        CK(cudaStreamWaitEvent(st, ctx->ev_copy_done, 0));  // the context stream ends after the last chunk
      else
        CK(cudaMemcpyAsync(K_dst, ctx->d_vals, sizeof(double) * (size_t)io_nnz(ctx), cudaMemcpyDefault, st));
      nvtxRangePop();
    }
    if (f_dst && f_dst != ctx->d_f)
      CK(cudaMemcpyAsync(f_dst, ctx->d_f, sizeof(double) * (size_t)io_rows(ctx), cudaMemcpyDefault, st));
    CK(cudaGetLastError());
    if (K_dst && stiff) ctx->last_d2h_bytes = 8 * (symmetric ? ctx->sym->n_packed : io_nnz(ctx));
    if (symmetric)
    {
      // host side of the block-upper transfer: as the packed rows of a phase arrive, the pool writes the full
      // CSR rows of that phase (lower blocks = transposes of upper blocks of earlier rows) while the next phases
      // are still being evaluated and copied
      nvtxRangePushA("host mirror");
      SymTransfer& T = *ctx->sym;
      const bool nt = ctx->nt_stores != 0;
      for (int k = 0; k < ctx->n_phases; ++k)
      {
        const int64_t n0 = T.phase_node[(size_t)k], n1 = T.phase_node[(size_t)k + 1];
        if (ctx->phase_off[(size_t)k + 1] <= ctx->phase_off[(size_t)k] || n1 <= n0) continue;
        CK(cudaEventSynchronize(T.ev_arrive[k]));
        const int64_t chunk = std::max<int64_t>(256, (n1 - n0 + 2 * T.n_threads - 1) / (2 * T.n_threads));
        for (int64_t r = n0; r < n1; r += chunk)
        {
          const int64_t r1 = std::min(n1, r + chunk);
          const SymTransfer* Tp = &T;
          T.pool.submit([Tp, r, r1, K_dst, nt] { sym_unpack_rows(*Tp, r, r1, K_dst, nt); });
        }
      }
      T.pool.wait();
      nvtxRangePop();
    }
    nvtxRangePop();
    return B200S_OK;
  }


  int b200s_set_row_export(b200s_ctx* ctx, int n_peers, const int32_t* peers, const int64_t* send_val, const int64_t* send_f,
      const int64_t* recv_offsets, const int64_t* recv_pos, const int64_t* recv_f_offsets, const int32_t* recv_fpos)
  {
    if (!ctx || n_peers < 0) return B200S_ERR_INVALID;
    if (ctx->n_rows_owned <= 0 || ctx->nnz_owned <= 0 || ctx->n_rows_owned > ctx->n_rows || ctx->nnz_owned > ctx->nnz)
      return fail(ctx, B200S_ERR_INVALID,
          "row export: options owned_rows / owned_nnz must be set before set_graph (rows %lld of %lld, nnz %lld of %lld)",
          (long long)ctx->n_rows_owned, (long long)ctx->n_rows, (long long)ctx->nnz_owned, (long long)ctx->nnz);
    cudaSetDevice(ctx->device);
    ctx->exp_peer.assign(peers, peers + n_peers);
    ctx->exp_send_val.assign(send_val, send_val + 2 * (size_t)n_peers);
    ctx->exp_send_f.assign(send_f, send_f + 2 * (size_t)n_peers);
    ctx->exp_recv_off.assign(recv_offsets, recv_offsets + n_peers + 1);
    ctx->exp_recv_foff.assign(recv_f_offsets, recv_f_offsets + n_peers + 1);
    for (int k = 0; k < n_peers; ++k)
    {
      const int64_t sv = send_val[2 * k], nv = send_val[2 * k + 1], sf = send_f[2 * k], nf = send_f[2 * k + 1];
      if (nv < 0 || nf < 0 || (nv > 0 && (sv < ctx->nnz_owned || sv + nv > ctx->nnz)) ||
          (nf > 0 && (sf < ctx->n_rows_owned || sf + nf > ctx->n_rows)))
        return fail(ctx, B200S_ERR_INVALID, "row export: send slice of peer %d lies outside the export rows", peers[k]);
    }
    const int64_t nr = ctx->exp_recv_off[(size_t)n_peers], nfr = ctx->exp_recv_foff[(size_t)n_peers];
    if (to_device(ctx, &ctx->d_exp_pos, recv_pos, nr) || to_device(ctx, &ctx->d_exp_fpos, recv_fpos, nfr) ||
        to_device(ctx, &ctx->d_exp_buf, (const double*)nullptr, nr) || to_device(ctx, &ctx->d_exp_fbuf, (const double*)nullptr, nfr))
      return B200S_ERR_CUDA;
    CK(cudaStreamSynchronize(ctx->stream));
    return B200S_OK;
  }

  int b200s_add_exported(b200s_ctx* ctx, int peer_index, const double* values, const double* f)
  {
    if (!ctx || peer_index < 0 || peer_index >= (int)ctx->exp_peer.size()) return B200S_ERR_INVALID;
    cudaSetDevice(ctx->device);
    const int64_t v0 = ctx->exp_recv_off[(size_t)peer_index], nv = ctx->exp_recv_off[(size_t)peer_index + 1] - v0;
    const int64_t f0 = ctx->exp_recv_foff[(size_t)peer_index], nf = ctx->exp_recv_foff[(size_t)peer_index + 1] - f0;
    if (values && nv > 0)
      CK(cudaMemcpyAsync(ctx->d_exp_buf + v0, values, sizeof(double) * (size_t)nv, cudaMemcpyDefault, ctx->stream));
    if (f && nf > 0) CK(cudaMemcpyAsync(ctx->d_exp_fbuf + f0, f, sizeof(double) * (size_t)nf, cudaMemcpyDefault, ctx->stream));
    if (!f)
    {
      // values only
      if (values && nv > 0)
      {
        k_add_exported<<<nblk(nv), 256, 0, ctx->stream>>>(nv, ctx->d_exp_pos + v0, ctx->d_exp_buf + v0, ctx->d_vals);
        ctx->launches += 1;
      }
    }
    else
      add_exported(ctx, peer_index, values != nullptr);
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaGetLastError());
    return B200S_OK;
  }

  int b200s_evaluate_async(b200s_ctx* ctx, int action)
  {
    if (!ctx) return B200S_ERR_INVALID;
    cudaSetDevice(ctx->device);
    ctx->timed_h2d = ctx->timed_d2h = false;  // no copies in this flavour: timers[4], [5] read 0
    return evaluate_impl(ctx, action, nullptr, nullptr);
  }

  int b200s_synchronize(b200s_ctx* ctx, b200s_error_info* err)
  {
    if (!ctx) return B200S_ERR_INVALID;
    cudaSetDevice(ctx->device);
    CK(cudaStreamSynchronize(ctx->stream));
    float ms = 0;
    auto el = [&](int a, int b)
    {
      ms = 0;
      cudaEventElapsedTime(&ms, ctx->ev[a], ctx->ev[b]);
      return (double)ms;
    };
    harvest_kernel_times(ctx);  // sets timers[2] = element kernel of the last evaluate
    ctx->timers[0] = el(1, 2), ctx->timers[1] = el(2, 3);
    ctx->timers[3] = el(1, 4);
    ctx->timers[4] = ctx->timed_h2d ? el(0, 1) : 0.0;
    ctx->timers[5] = ctx->timed_d2h ? el(4, 5) : 0.0;
    int code = B200S_OK;
    const ErrWord& w = *ctx->h_err;
    if (w.n_neg_jac > 0)
      code = fail(ctx, B200S_ERR_NEG_JACOBIAN,
          "determinant of jacobian is <= 0 at one node of the element (column element %lld, %llu "
          "elements affected, min det J %g)",
          w.first_bad, w.n_neg_jac, w.min_detJ);
    else if (ctx->check_finite && w.n_nonfinite > 0)
      code = fail(ctx, B200S_ERR_NONFINITE, "%llu non-finite entries in f/K (min det F %g)",
          w.n_nonfinite, w.min_detF);
    if (err)
    {
      err->code = code;
      err->n_neg_jacobian = (int32_t)(w.n_neg_jac > INT32_MAX ? INT32_MAX : w.n_neg_jac);
      err->first_bad_element = w.n_neg_jac ? w.first_bad : -1;
      err->min_detJ = w.min_detJ, err->min_detF = w.min_detF;
      err->n_nonfinite = (int64_t)w.n_nonfinite;
    }
    return code;
  }

  int b200s_evaluate(b200s_ctx* ctx, int action, const double* u_row, double* K_values,
      double* f_row, b200s_error_info* err)
  {
    if (!ctx) return B200S_ERR_INVALID;
    if (!ctx->d_pos) return fail(ctx, B200S_ERR_INVALID, "set_graph must precede evaluate");
    cudaSetDevice(ctx->device);
    cudaStream_t st = ctx->stream;
    ctx->timed_h2d = u_row != nullptr, ctx->timed_d2h = (K_values || f_row);
    CK(cudaEventRecord(ctx->ev[0], st));
    if (u_row && u_row != ctx->d_u_row)
    {
      nvtxRangePushA("h2d");
      CK(cudaMemcpyAsync(ctx->d_u_row, u_row, sizeof(double) * (size_t)io_rows(ctx), cudaMemcpyDefault, st));
      nvtxRangePop();
    }
    const int rc = evaluate_impl(ctx, action, K_values, f_row);
    if (rc) return rc;
    CK(cudaEventRecord(ctx->ev[5], st));
    return b200s_synchronize(ctx, err);
  }

  double* b200s_device_values(b200s_ctx* ctx) { return ctx ? ctx->d_vals : nullptr; }
  double* b200s_device_f(b200s_ctx* ctx) { return ctx ? ctx->d_f : nullptr; }
  double* b200s_device_u_row(b200s_ctx* ctx) { return ctx ? ctx->d_u_row : nullptr; }
  double* b200s_device_u_col(b200s_ctx* ctx) { return ctx ? ctx->d_u_col : nullptr; }
  void* b200s_stream(b200s_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
  int64_t b200s_nnz(const b200s_ctx* ctx) { return ctx ? ctx->nnz : 0; }
  int64_t b200s_n_rows(const b200s_ctx* ctx) { return ctx ? ctx->n_rows : 0; }

  int b200s_get_timers(b200s_ctx* ctx, double* out6)
  {
    if (!ctx || !out6) return B200S_ERR_INVALID;
    for (int i = 0; i < 6; ++i) out6[i] = ctx->timers[i];
    return B200S_OK;
  }
  int64_t b200s_launch_count(const b200s_ctx* ctx) { return ctx ? ctx->launches : 0; }

  int b200s_kernel_time(b200s_ctx* ctx, double* sum_ms, int64_t* count, int reset)
  {
    if (!ctx) return B200S_ERR_INVALID;
    if (sum_ms) *sum_ms = ctx->ktime_sum_ms;
    if (count) *count = ctx->ktime_count;
    if (reset) ctx->ktime_sum_ms = 0.0, ctx->ktime_count = 0;
    return B200S_OK;
  }

  // ---- further materials / states -------------------------------------------------------------------
  int b200s_set_material_elasthyper(b200s_ctx* ctx, int n_summands, const int32_t* kinds,
      const int32_t* modes, const double* params)
  {
    if (!ctx || !kinds || !params || n_summands < 1) return B200S_ERR_INVALID;
    if (n_summands > HYPER_MAX_SUMMANDS)
      return fail(ctx, B200S_ERR_UNSUPPORTED, "at most %d summands on the device path", HYPER_MAX_SUMMANDS);
    cudaSetDevice(ctx->device);
    HyperMat hm{};
    hm.n = n_summands, hm.uses_modified = 0;
    for (int k = 0; k < n_summands; ++k)
    {
      if (summand_num_params(kinds[k]) < 0)
        return fail(ctx, B200S_ERR_UNSUPPORTED, "summand kind %d is not on the device path", kinds[k]);
      if (!make_summand(kinds[k], params + 4 * k, modes ? modes[k] : 0, hm.s[k]))
        return fail(ctx, B200S_ERR_INVALID, "summand %d (kind %d): parameters rejected", k, kinds[k]);
      if (summand_is_modified(kinds[k])) hm.uses_modified = 1;
    }
    ctx->hm = hm;
    if (to_device(ctx, &ctx->d_hyper, &ctx->hm, 1)) return B200S_ERR_CUDA;
    CK(cudaStreamSynchronize(ctx->stream));
    MatParams mp{};
    for (double& v : mp.p) v = 0.0;
    mp.hyper = ctx->d_hyper;
    ctx->matkind = B200S_MAT_ELASTHYPER, ctx->mp = mp;
    ctx->general_tangent = false;
    dfree(ctx->d_per_ele), dfree(ctx->d_mat_hist);
    return B200S_OK;
  }

  int b200s_set_element_fibers(b200s_ctx* ctx, const double* fiber1)
  {
    if (!ctx) return B200S_ERR_INVALID;
    if (ctx->celltype < 0) return fail(ctx, B200S_ERR_INVALID, "set_mesh must precede set_element_fibers");
    cudaSetDevice(ctx->device);
    dfree(ctx->d_fibers);
    if (!fiber1) return B200S_OK;
    // read_fibers normalises the input vectors (src/solid_ele/4C_solid_ele_fiber.cpp:66-69)
    const int64_t n = ctx->n_col_eles;
    std::vector<double> h(3 * (size_t)(n > 0 ? n : 1));
    CK(cudaMemcpy(h.data(), fiber1, sizeof(double) * 3 * (size_t)n, cudaMemcpyDefault));
    for (int64_t e = 0; e < n; ++e)
    {
      const double nrm = std::sqrt(h[3 * e] * h[3 * e] + h[3 * e + 1] * h[3 * e + 1] + h[3 * e + 2] * h[3 * e + 2]);
      if (!(nrm > 0.0)) return fail(ctx, B200S_ERR_INVALID, "element %lld has a zero fibre vector", (long long)e);
      for (int d = 0; d < 3; ++d) h[3 * e + d] /= nrm;
    }
    double* tmp = nullptr;
    if (to_device(ctx, &tmp, h.data(), 3 * n)) return B200S_ERR_CUDA;
    if (to_device(ctx, &ctx->d_fibers, (const double*)nullptr, 3 * n)) return B200S_ERR_CUDA;
    if (n > 0)
    {
      // caller's element order -> device (locality) order
      k_hist_permute<<<nblk(3 * n), 256, 0, ctx->stream>>>(n, 3, ctx->d_ele_id, tmp, ctx->d_fibers, 0);
      ctx->launches += 1;
    }
    CK(cudaStreamSynchronize(ctx->stream));
    cudaFree(tmp);
    CK(cudaGetLastError());
    return B200S_OK;
  }

  int b200s_set_density(b200s_ctx* ctx, double density)
  {
    if (!ctx) return B200S_ERR_INVALID;
    if (!(density >= 0.0)) return fail(ctx, B200S_ERR_INVALID, "density must be >= 0");
    ctx->density = density;
    return B200S_OK;
  }

  int b200s_set_time_step_size(b200s_ctx* ctx, double dt)
  {
    if (!ctx) return B200S_ERR_INVALID;
    if (!(dt >= 0.0)) return fail(ctx, B200S_ERR_INVALID, "time step size must be >= 0");
    ctx->dt = dt, ctx->dt_set = true;
    if (ctx->matkind == B200S_MAT_VISCONEOHOOKE) make_visco_neohooke(ctx->visco_in, dt, ctx->mp.p);
    if (ctx->matkind == B200S_MAT_PLASTIC_NLNLOG && ctx->d_mat_ext)
    {
      double q[PLASTIC_NPAR];
      if (!make_plastic_nlnlog(ctx->plastic_in, dt, q)) return fail(ctx, B200S_ERR_INVALID, "time step size must be > 0 for MAT_Struct_PlasticNlnLogNeoHooke");
      cudaSetDevice(ctx->device);
      CK(cudaMemcpyAsync(ctx->d_mat_ext, q, sizeof q, cudaMemcpyHostToDevice, ctx->stream));
      CK(cudaStreamSynchronize(ctx->stream));  // q is a local
    }
    return B200S_OK;
  }

  int b200s_set_row_elements(b200s_ctx* ctx, const uint8_t* is_row)
  {
    if (!ctx) return B200S_ERR_INVALID;
    if (ctx->celltype < 0) return fail(ctx, B200S_ERR_INVALID, "set_mesh must precede set_row_elements");
    cudaSetDevice(ctx->device);
    dfree(ctx->d_ele_is_row);
    if (!is_row) return B200S_OK;
    if (to_device(ctx, &ctx->d_ele_is_row, is_row, ctx->n_col_eles)) return B200S_ERR_CUDA;
    CK(cudaStreamSynchronize(ctx->stream));
    return B200S_OK;
  }

  int b200s_set_state(b200s_ctx* ctx, const char* name, const double* v_row)
  {
    if (!ctx || !name || !v_row) return B200S_ERR_INVALID;
    if (!ctx->d_rowptr) return fail(ctx, B200S_ERR_INVALID, "set_graph must precede set_state");
    cudaSetDevice(ctx->device);
    double* dst = nullptr;
    if (!strcmp(name, "displacement"))
      dst = ctx->d_u_row;
    else if (!strcmp(name, "acceleration"))
    {
      if (!ctx->d_acc_row && to_device(ctx, &ctx->d_acc_row, (const double*)nullptr, ctx->n_rows))
        return B200S_ERR_CUDA;
      dst = ctx->d_acc_row;
    }
    else
      return fail(ctx, B200S_ERR_UNSUPPORTED, "state '%s' is not used on the device path", name);
    if (v_row != dst)
      CK(cudaMemcpyAsync(dst, v_row, sizeof(double) * (size_t)io_rows(ctx), cudaMemcpyDefault, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return B200S_OK;
  }

  double* b200s_device_mass_values(b200s_ctx* ctx) { return ctx ? ctx->d_mass : nullptr; }
  int b200s_get_values(b200s_ctx* ctx, int which, double* values_out)
  {
    if (!ctx || !values_out) return B200S_ERR_INVALID;
    const double* src = which == B200S_MATRIX_MASS ? ctx->d_mass : ctx->d_vals;
    if (!src || ctx->nnz <= 0) return fail(ctx, B200S_ERR_INVALID, "matrix %d has not been assembled", which);
    cudaSetDevice(ctx->device);
    CK(cudaMemcpyAsync(values_out, src, sizeof(double) * (size_t)ctx->nnz, cudaMemcpyDefault, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return B200S_OK;
  }
  double* b200s_device_f_inertia(b200s_ctx* ctx) { return ctx ? ctx->d_f_inertia : nullptr; }

  // rows [row0,row1) of the resident CSR: what ExtractMyRowView would show for those rows
  int b200s_get_value_rows(b200s_ctx* ctx, int which, int64_t row0, int64_t row1, double* values_out,
      int32_t* colind_out, int64_t* rowptr_out)
  {
    if (!ctx || row0 < 0 || row1 < row0) return B200S_ERR_INVALID;
    if (!ctx->d_rowptr || row1 > ctx->n_rows) return fail(ctx, B200S_ERR_INVALID, "rows [%lld,%lld) outside the graph",
        (long long)row0, (long long)row1);
    cudaSetDevice(ctx->device);
    const int64_t nr = row1 - row0;
    std::vector<int64_t> rp((size_t)nr + 1);
    CK(cudaMemcpyAsync(rp.data(), ctx->d_rowptr + row0, sizeof(int64_t) * (size_t)(nr + 1), cudaMemcpyDeviceToHost,
        ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    const int64_t lo = rp[0], n = rp[(size_t)nr] - lo;
    if (rowptr_out)
    {
      for (auto& v : rp) v -= lo;  // offsets relative to the first returned entry
      CK(cudaMemcpyAsync(rowptr_out, rp.data(), sizeof(int64_t) * (size_t)(nr + 1), cudaMemcpyDefault, ctx->stream));
    }
    if (values_out && n > 0)
    {
      const double* src = matrix_values(ctx, which);
      if (!src) return fail(ctx, B200S_ERR_INVALID, "matrix %d has not been assembled", which);
      CK(cudaMemcpyAsync(values_out, src + lo, sizeof(double) * (size_t)n, cudaMemcpyDefault, ctx->stream));
    }
    if (colind_out && n > 0)
      CK(cudaMemcpyAsync(colind_out, ctx->d_colind + lo, sizeof(int32_t) * (size_t)n, cudaMemcpyDefault, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return B200S_OK;
  }
  int b200s_get_f_rows(b200s_ctx* ctx, int64_t row0, int64_t row1, double* f_out)
  {
    if (!ctx || !f_out || row0 < 0 || row1 < row0) return B200S_ERR_INVALID;
    if (!ctx->d_f || row1 > ctx->n_rows) return fail(ctx, B200S_ERR_INVALID, "rows outside the residual vector");
    cudaSetDevice(ctx->device);
    CK(cudaMemcpyAsync(f_out, ctx->d_f + row0, sizeof(double) * (size_t)(row1 - row0), cudaMemcpyDefault, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return B200S_OK;
  }

  int b200s_get_mass(b200s_ctx* ctx, double* M_values, double* f_inertia)
  {
    if (!ctx) return B200S_ERR_INVALID;
    cudaSetDevice(ctx->device);
    if (M_values)
    {
      if (!ctx->d_mass) return fail(ctx, B200S_ERR_INVALID, "no mass matrix evaluated yet");
      CK(cudaMemcpyAsync(M_values, ctx->d_mass, sizeof(double) * (size_t)ctx->nnz, cudaMemcpyDefault, ctx->stream));
    }
    if (f_inertia)
    {
      if (!ctx->d_f_inertia) return fail(ctx, B200S_ERR_INVALID, "no inertia force evaluated yet");
      CK(cudaMemcpyAsync(f_inertia, ctx->d_f_inertia, sizeof(double) * (size_t)ctx->n_rows, cudaMemcpyDefault,
          ctx->stream));
    }
    CK(cudaStreamSynchronize(ctx->stream));
    return B200S_OK;
  }

  // ---- struct_calc_stress / struct_calc_energy --------------------------------------------------------
  static int launch_post(b200s_ctx* ctx, int stress_type, int strain_type, double* d_stress,
      double* d_strain, double* d_energy)
  {
    if (ctx->celltype < 0 || ctx->matkind < 0)
      return fail(ctx, B200S_ERR_INVALID, "set_mesh and set_material must precede this action");
    if (material_needs_fibers(ctx) && !ctx->d_fibers)
      return fail(ctx, B200S_ERR_INVALID, "the material has a fibre summand: b200s_set_element_fibers first");
    const bool ans = ctx->tech == B200S_TECH_SHELL_ANS || ctx->tech == B200S_TECH_SHELL_EAS_ANS;
    if (ans && (ctx->celltype != B200S_HEX8 || !ctx->d_dN_ans))
      return fail(ctx, B200S_ERR_UNSUPPORTED, "TECH shell_ans / shell_eas_ans on the device: hex8 blocks");
    PostArgs P;
    P.dN_ans = ans ? ctx->d_dN_ans : nullptr;
    P.X = ctx->dX, P.u = ctx->d_u_col, P.conn = ctx->dconn, P.per_ele_param = ctx->d_per_ele;
    P.fibers = ctx->d_fibers;
    P.ele_id = ctx->d_ele_id, P.ele_is_row = ctx->d_ele_is_row;
    P.dN_gp = ctx->d_dN_gp, P.dN_cen = ctx->d_dN_cen, P.w = ctx->d_w, P.mp = ctx->mp;
    P.gp_hist = ctx->prestress == B200S_PRESTRESS_MULF ? ctx->d_gp_hist : nullptr;
    P.hist_pts = mulf_points(ctx);
    P.mat_hist = ctx->d_mat_hist;
    P.eas_neas = eas_neas(ctx->tech);
    P.eas_hist_size = P.eas_neas ? hex8eas::hist_size(P.eas_neas) : 0;
    P.eas_t0_stride = hex8eas::T0_STRIDE;
    P.eas_hist = ctx->d_eas_hist, P.eas_t0 = ctx->d_eas_t0, P.xi_gp = ctx->d_xi_gp;
    if (P.eas_neas && (!ctx->d_eas_hist || !ctx->d_eas_t0))
      return fail(ctx, B200S_ERR_INVALID, "EAS history missing");
    P.n_ele = ctx->n_col_eles, P.nn = ctx->nn, P.ngp = ctx->ngp;
    P.fbar = ctx->tech == B200S_TECH_FBAR;
    P.linear = ctx->kinem == B200S_KINEM_LINEAR;
    P.stress_type = stress_type, P.strain_type = strain_type;
    P.stress = d_stress, P.strain = d_strain, P.energy = d_energy;
    const unsigned grid = nblk(ctx->n_col_eles * ctx->ngp, 128);
    if (!grid) return B200S_OK;
    switch (ctx->matkind)
    {
      case B200S_MAT_COUP_NEOHOOKE:
        k_post_points<MAT_COUP_NEOHOOKE><<<grid, 128, 0, ctx->stream>>>(P);
        break;
      case B200S_MAT_ISO_VOL:
        k_post_points<MAT_ISO_VOL><<<grid, 128, 0, ctx->stream>>>(P);
        break;
      case B200S_MAT_SVK:
        k_post_points<MAT_SVK><<<grid, 128, 0, ctx->stream>>>(P);
        break;
      case B200S_MAT_ELASTHYPER:
        k_post_points<MAT_HYPER><<<grid, 128, 0, ctx->stream>>>(P);
        break;
      case B200S_MAT_VISCONEOHOOKE:
        if (d_energy) return fail(ctx, B200S_ERR_UNSUPPORTED, "MAT_VISCONEOHOOKE has no strain energy");
        if (!ctx->dt_set) return fail(ctx, B200S_ERR_INVALID, "Time step size not given in evaluation context.");
        k_post_points<MAT_VISCO><<<grid, 128, 0, ctx->stream>>>(P);
        break;
      case B200S_MAT_PLASTIC_NLNLOG:
        if (d_energy) return fail(ctx, B200S_ERR_UNSUPPORTED, "MAT_Struct_PlasticNlnLogNeoHooke has no strain energy");
        if (!ctx->dt_set) return fail(ctx, B200S_ERR_INVALID, "Time step size not given in evaluation context.");
        k_post_points<MAT_PLASTIC><<<grid, 128, 0, ctx->stream>>>(P);
        break;
      case B200S_MAT_DRUCKER_PRAGER:
        if (d_energy) return fail(ctx, B200S_ERR_UNSUPPORTED, "MAT_Struct_DruckerPrager has no strain energy");
        if (!P.linear) return fail(ctx, B200S_ERR_UNSUPPORTED, "MAT_Struct_DruckerPrager needs KINEM linear");
        k_post_points<MAT_DP><<<grid, 128, 0, ctx->stream>>>(P);
        break;
      default:
        return fail(ctx, B200S_ERR_INVALID, "material not set");
    }
    ctx->launches += 1;
    return B200S_OK;
  }

  int b200s_calc_stress(b200s_ctx* ctx, int stress_type, int strain_type, double* stress_out,
      double* strain_out)
  {
    if (!ctx) return B200S_ERR_INVALID;
    if (stress_type < B200S_STRESS_NONE || stress_type > B200S_STRESS_CAUCHY ||
        strain_type < B200S_STRAIN_NONE || strain_type > B200S_STRAIN_EA)
      return fail(ctx, B200S_ERR_UNSUPPORTED, "stress/strain output type %d/%d has no device kernel",
          stress_type, strain_type);
    if (!ctx->d_rowptr) return fail(ctx, B200S_ERR_INVALID, "set_graph must precede calc_stress");
    cudaSetDevice(ctx->device);
    const size_t n = (size_t)ctx->n_col_eles * ctx->ngp * 6;
    double *d_s = nullptr, *d_e = nullptr;
    const bool want_s = stress_out && stress_type != B200S_STRESS_NONE;
    const bool want_e = strain_out && strain_type != B200S_STRAIN_NONE;
    if (want_s) CK(cudaMalloc((void**)&d_s, sizeof(double) * (n ? n : 1)));
    if (want_e) CK(cudaMalloc((void**)&d_e, sizeof(double) * (n ? n : 1)));
    int rc = import_row_to_col(ctx, ctx->d_u_row, ctx->d_u_col, false);
    if (!rc) rc = launch_post(ctx, stress_type, strain_type, d_s, d_e, nullptr);
    if (!rc && want_s && cudaMemcpyAsync(stress_out, d_s, sizeof(double) * n, cudaMemcpyDefault, ctx->stream))
      rc = B200S_ERR_CUDA;
    if (!rc && want_e && cudaMemcpyAsync(strain_out, d_e, sizeof(double) * n, cudaMemcpyDefault, ctx->stream))
      rc = B200S_ERR_CUDA;
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess && !rc)
      rc = fail(ctx, B200S_ERR_CUDA, "calc_stress: %s", cudaGetErrorString(cudaGetLastError()));
    cudaFree(d_s), cudaFree(d_e);
    return rc;
  }

  int b200s_calc_energy(b200s_ctx* ctx, double* energy_out)
  {
    if (!ctx || !energy_out) return B200S_ERR_INVALID;
    if (!ctx->d_rowptr) return fail(ctx, B200S_ERR_INVALID, "set_graph must precede calc_energy");
    if (ctx->nranks > 1 && !ctx->d_ele_is_row)
      return fail(ctx, B200S_ERR_INVALID,
          "calc_energy on a partitioned mesh needs b200s_set_row_elements (row elements only)");
    if (ctx->matkind == B200S_MAT_ELASTHYPER)
      for (int k = 0; k < ctx->hm.n; ++k)
        if (summand_in_stretches(ctx->hm.s[k].kind))
          return fail(ctx, B200S_ERR_UNSUPPORTED, "Summand does not support calculation of strain energy");
    cudaSetDevice(ctx->device);
    if (!ctx->d_energy && to_device(ctx, &ctx->d_energy, (const double*)nullptr, 1)) return B200S_ERR_CUDA;
    CK(cudaMemsetAsync(ctx->d_energy, 0, sizeof(double), ctx->stream));
    int rc = import_row_to_col(ctx, ctx->d_u_row, ctx->d_u_col, false);
    if (!rc) rc = launch_post(ctx, B200S_STRESS_NONE, B200S_STRAIN_NONE, nullptr, nullptr, ctx->d_energy);
    if (rc) return rc;
    CK(cudaMemcpyAsync(energy_out, ctx->d_energy, sizeof(double), cudaMemcpyDefault, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaGetLastError());
    return B200S_OK;
  }

  // ---- MULF prestress: struct_update_prestress and the Gauss-point history (restart) ------------------
  int b200s_update_prestress(b200s_ctx* ctx)
  {
    if (!ctx) return B200S_ERR_INVALID;
    if (ctx->prestress != B200S_PRESTRESS_MULF || !ctx->d_gp_hist)
      return fail(ctx, B200S_ERR_UNSUPPORTED,
          "Your element evaluation does not allow to update prestress. You may need to add MULF to your "
          "element line definitions.");
    if (!ctx->d_rowptr) return fail(ctx, B200S_ERR_INVALID, "set_graph must precede update_prestress");
    cudaSetDevice(ctx->device);
    const int rc = import_row_to_col(ctx, ctx->d_u_row, ctx->d_u_col, false);
    if (rc) return rc;
    // MulfFBarFormulation::update_prestress: the centroid history is updated like a Gauss point's (mulf_fbar.hpp:245-277)
    const int64_t n = ctx->n_col_eles * mulf_points(ctx);
    if (n > 0)
    {
      k_mulf_update<<<nblk(n, 128), 128, 0, ctx->stream>>>(ctx->n_col_eles, ctx->nn, mulf_points(ctx), ctx->dconn, ctx->dX,
          ctx->d_u_col, ctx->d_dN_pts, ctx->d_gp_hist);
      ctx->launches += 1;
    }
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaGetLastError());
    return B200S_OK;
  }

  int64_t b200s_gp_history_size(const b200s_ctx* ctx)
  {
    if (!ctx || ctx->prestress != B200S_PRESTRESS_MULF) return 0;
    return ctx->n_col_eles * mulf_points(ctx) * MULF_HIST;
  }

  // ---- EAS: struct_calc_recover and the element history ---------------------------------------------
  int b200s_recover(b200s_ctx* ctx, const double* du_row, double step_length)
  {
    if (!ctx || !du_row) return B200S_ERR_INVALID;
    if (!ctx->d_rowptr) return fail(ctx, B200S_ERR_INVALID, "set_graph must precede recover");
    const int neas = eas_neas(ctx->tech);
    if (!neas) return B200S_OK;  // formulations without condensed variables: nothing to recover
    if (!ctx->d_eas_hist) return fail(ctx, B200S_ERR_INVALID, "EAS history missing");
    cudaSetDevice(ctx->device);
    if (!ctx->d_du_row && to_device(ctx, &ctx->d_du_row, (const double*)nullptr, ctx->n_rows)) return B200S_ERR_CUDA;
    if (!ctx->d_du_col && to_device(ctx, &ctx->d_du_col, (const double*)nullptr, 3 * ctx->n_col_nodes))
      return B200S_ERR_CUDA;
    if (du_row != ctx->d_du_row)
      CK(cudaMemcpyAsync(ctx->d_du_row, du_row, sizeof(double) * (size_t)ctx->n_rows, cudaMemcpyDefault, ctx->stream));
    const int rc = import_row_to_col(ctx, ctx->d_du_row, ctx->d_du_col, false);
    if (rc) return rc;
    if (b200s_parts::eas_recover(ctx, neas, step_length)) return B200S_ERR_CUDA;
    ctx->eas_old_step_length = step_length;
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaGetLastError());
    return B200S_OK;
  }

  // correct_condensed_variables_for_linesearch (4C_solid_ele_calc_eas.hpp:228-243; correct_alpha,
  // eas_helpers.hpp:192-199): alpha += (new - old step length) * alpha_inc
  int b200s_recover_linesearch(b200s_ctx* ctx, double new_step_length)
  {
    if (!ctx) return B200S_ERR_INVALID;
    const int neas = eas_neas(ctx->tech);
    if (!neas) return B200S_OK;
    if (!ctx->d_eas_hist) return fail(ctx, B200S_ERR_INVALID, "EAS history missing");
    cudaSetDevice(ctx->device);
    const int hs = hex8eas::hist_size(neas);
    k_eas_correct_alpha<<<nblk(ctx->n_col_eles * neas), 256, 0, ctx->stream>>>(ctx->n_col_eles, neas, hs,
        new_step_length - ctx->eas_old_step_length, ctx->d_eas_hist);
    ctx->launches += 1;
    ctx->eas_old_step_length = new_step_length;
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaGetLastError());
    return B200S_OK;
  }

  // struct_calc_update_istep / struct_calc_reset_istep (SolidEleCalc::update :317-363 and
  // reset_to_last_converged :570-575 of 4C_solid_ele_calc.cpp): old_history_data_ = history_data_ and back.
  // The materials on the device path have no history of their own (So3Material::update / reset_step are
  // no-ops for them), so only the element (EAS) / Gauss-point (MULF) history is involved.
  static int history_snapshot(b200s_ctx* ctx, int restore)
  {
    if (!ctx) return B200S_ERR_INVALID;
    double* cur = eas_neas(ctx->tech) ? ctx->d_eas_hist : (ctx->prestress == B200S_PRESTRESS_MULF ? ctx->d_gp_hist : nullptr);
    if (!cur) return B200S_OK;  // formulations without history: nothing to do
    const int64_t n = eas_neas(ctx->tech) ? ctx->n_col_eles * hex8eas::hist_size(eas_neas(ctx->tech))
                                          : ctx->n_col_eles * mulf_points(ctx) * MULF_HIST;
    cudaSetDevice(ctx->device);
    if (!ctx->d_hist_old)
    {
      if (restore)
      {
        // no converged step yet: old_history_data_ is still the default-constructed history
        // (4C_solid_ele_calc.cpp:570-575), i.e. zero EAS parameters / the reference-configuration MULF history
        if (eas_neas(ctx->tech))
        {
          CK(cudaMemsetAsync(cur, 0, sizeof(double) * (size_t)(n > 0 ? n : 1), ctx->stream));
          ctx->eas_old_step_length = 1.0;
          CK(cudaStreamSynchronize(ctx->stream));
          return B200S_OK;
        }
        return mulf_allocate(ctx) ? B200S_ERR_CUDA : B200S_OK;
      }
      if (to_device(ctx, &ctx->d_hist_old, (const double*)nullptr, n)) return B200S_ERR_CUDA;
    }
    CK(cudaMemcpyAsync(restore ? cur : ctx->d_hist_old, restore ? ctx->d_hist_old : cur, sizeof(double) * (size_t)n,
        cudaMemcpyDeviceToDevice, ctx->stream));
    if (restore)
      ctx->eas_old_step_length = ctx->eas_old_step_length_snap;
    else
      ctx->eas_old_step_length_snap = ctx->eas_old_step_length;
    CK(cudaStreamSynchronize(ctx->stream));
    return B200S_OK;
  }
  // history material: last <- current (update) or current <- last (reset), in place
  static int material_history_commit(b200s_ctx* ctx, int restore)
  {
    if (!ctx->d_mat_hist) return B200S_OK;
    cudaSetDevice(ctx->device);
    const int64_t np = ctx->n_col_eles * ctx->ngp;
    if (np > 0)
    {
      // update: last <- current (Mat::ViscoNeoHooke::update additionally clears the current iterate, :197-216);
      // reset: current <- last
      const int H = mat_hist_per_gp(ctx);
      k_material_history_commit<<<nblk(np * (H / 2)), 256, 0, ctx->stream>>>(np, H, H / 2,
          restore ? 1 : (ctx->matkind == B200S_MAT_VISCONEOHOOKE ? 2 : (ctx->matkind == B200S_MAT_PLASTIC_NLNLOG ? 3 : 0)),
          ctx->d_mat_hist);
      ctx->launches += 1;
    }
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaGetLastError());
    return B200S_OK;
  }
  int b200s_update_step(b200s_ctx* ctx)
  {
    if (!ctx) return B200S_ERR_INVALID;
    const int rc = material_history_commit(ctx, 0);
    return rc ? rc : history_snapshot(ctx, 0);
  }
  int b200s_reset_step(b200s_ctx* ctx)
  {
    if (!ctx) return B200S_ERR_INVALID;
    const int rc = material_history_commit(ctx, 1);
    return rc ? rc : history_snapshot(ctx, 1);
  }
  int64_t b200s_material_history_size(const b200s_ctx* ctx)
  {
    return (ctx && ctx->d_mat_hist) ? ctx->n_col_eles * ctx->ngp * mat_hist_per_gp(ctx) : 0;
  }

  int64_t b200s_element_history_size(const b200s_ctx* ctx)
  {
    if (!ctx || !eas_neas(ctx->tech)) return 0;
    return ctx->n_col_eles * hex8eas::hist_size(eas_neas(ctx->tech));
  }

  static int history_copy(b200s_ctx* ctx, double* caller, int to_caller, int element_level = 0)
  {
    if (!ctx || !caller) return B200S_ERR_INVALID;
    // element_level: 0 MULF Gauss-point history, 1 EAS element history, 2 material history
    double* d_hist = element_level == 2 ? ctx->d_mat_hist : (element_level ? ctx->d_eas_hist : ctx->d_gp_hist);
    if (!d_hist || (element_level == 2 ? false
                                       : (element_level ? !eas_neas(ctx->tech) : ctx->prestress != B200S_PRESTRESS_MULF)))
      return fail(ctx, B200S_ERR_INVALID, "the element block has no %s history",
          element_level == 2 ? "material" : (element_level ? "element" : "Gauss-point"));
    cudaSetDevice(ctx->device);
    const int64_t n = element_level == 2 ? b200s_material_history_size(ctx)
                                         : (element_level ? b200s_element_history_size(ctx) : b200s_gp_history_size(ctx));
    const int per_ele = element_level == 2 ? ctx->ngp * mat_hist_per_gp(ctx)
                                           : (element_level ? hex8eas::hist_size(eas_neas(ctx->tech)) : mulf_points(ctx) * MULF_HIST);
    double* tmp = nullptr;
    CK(cudaMalloc((void**)&tmp, sizeof(double) * (size_t)(n ? n : 1)));
    int rc = B200S_OK;
    if (to_caller)
    {
      k_hist_permute<<<nblk(n), 256, 0, ctx->stream>>>(ctx->n_col_eles, per_ele, ctx->d_ele_id, d_hist, tmp, 1);
      if (cudaMemcpyAsync(caller, tmp, sizeof(double) * (size_t)n, cudaMemcpyDefault, ctx->stream)) rc = B200S_ERR_CUDA;
    }
    else
    {
      if (cudaMemcpyAsync(tmp, caller, sizeof(double) * (size_t)n, cudaMemcpyDefault, ctx->stream)) rc = B200S_ERR_CUDA;
      k_hist_permute<<<nblk(n), 256, 0, ctx->stream>>>(ctx->n_col_eles, per_ele, ctx->d_ele_id, tmp, d_hist, 0);
    }
    ctx->launches += 1;
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess || cudaGetLastError() != cudaSuccess) rc = B200S_ERR_CUDA;
    cudaFree(tmp);
    return rc ? fail(ctx, rc, "history copy failed") : B200S_OK;
  }
  double b200s_eas_old_step_length(const b200s_ctx* ctx) { return ctx ? ctx->eas_old_step_length : 1.0; }
  int b200s_set_eas_old_step_length(b200s_ctx* ctx, double step_length)
  {
    if (!ctx) return B200S_ERR_INVALID;
    ctx->eas_old_step_length = step_length;
    return B200S_OK;
  }
  int b200s_get_gp_history(b200s_ctx* ctx, double* hist_out) { return history_copy(ctx, hist_out, 1); }
  int b200s_set_gp_history(b200s_ctx* ctx, const double* hist_in)
  {
    return history_copy(ctx, const_cast<double*>(hist_in), 0);
  }
  int b200s_get_material_history(b200s_ctx* ctx, double* hist_out) { return history_copy(ctx, hist_out, 1, 2); }
  int b200s_set_material_history(b200s_ctx* ctx, const double* hist_in)
  {
    return history_copy(ctx, const_cast<double*>(hist_in), 0, 2);
  }
  int b200s_get_element_history(b200s_ctx* ctx, double* hist_out) { return history_copy(ctx, hist_out, 1, 1); }
  int b200s_set_element_history(b200s_ctx* ctx, const double* hist_in)
  {
    return history_copy(ctx, const_cast<double*>(hist_in), 0, 1);
  }

  // ---- the step after the path: Dirichlet conditions and products on the device ---------------------
  // stages a row-map vector that may live on the host
  static int stage_row_vector(b200s_ctx* ctx, double** slot, const double* src)
  {
    if (!*slot && to_device(ctx, slot, (const double*)nullptr, ctx->n_rows)) return B200S_ERR_CUDA;
    if (src != *slot)
      CK(cudaMemcpyAsync(*slot, src, sizeof(double) * (size_t)ctx->n_rows, cudaMemcpyDefault, ctx->stream));
    return B200S_OK;
  }
  static double* matrix_values(b200s_ctx* ctx, int which)
  {
    return which == B200S_MATRIX_MASS ? ctx->d_mass : ctx->d_vals;
  }

  // ---- external loads (Discretization::evaluate_neumann on the element block) -----------------------
  int b200s_clear_neumann(b200s_ctx* ctx)
  {
    if (!ctx) return B200S_ERR_INVALID;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (auto& c : ctx->neumann) dfree(c.d_faces), dfree(c.d_ele_slot), dfree(c.d_gp_fac);
    ctx->neumann.clear();
    return B200S_OK;
  }

  int b200s_add_surface_neumann(b200s_ctx* ctx, int neumann_type, int64_t n_faces, const int32_t* face_nodes,
      const int32_t* onoff, const double* val)
  {
    if (!ctx || n_faces < 0 || (n_faces > 0 && !face_nodes) || !onoff || !val) return B200S_ERR_INVALID;
    if (ctx->celltype < 0) return fail(ctx, B200S_ERR_INVALID, "set_mesh must precede add_surface_neumann");
    if (ctx->face_nn == 0)
      return fail(ctx, B200S_ERR_UNSUPPORTED, "surface loads on wedge6 / pyramid5 / hex18 blocks (two surface shapes) have no device kernel");
    if (neumann_type < B200S_NEUMANN_LIVE || neumann_type > B200S_NEUMANN_ORTHOPRESSURE)
      return fail(ctx, B200S_ERR_UNSUPPORTED, "Unknown type of SurfaceNeumann condition");
    if (neumann_type != B200S_NEUMANN_LIVE && (onoff[0] != 1 || onoff[1] != 0 || onoff[2] != 0))
      return fail(ctx, B200S_ERR_INVALID, "orthopressure on 1st dof only!");
    for (int64_t t = 0; t < n_faces * ctx->face_nn; ++t)
      if (face_nodes[t] < 0 || face_nodes[t] >= ctx->n_col_nodes)
        return fail(ctx, B200S_ERR_INVALID, "face %lld: node id %d out of range", (long long)(t / ctx->face_nn),
            face_nodes[t]);
    cudaSetDevice(ctx->device);
    b200s_ctx::NeumannCond c;
    c.volume = false, c.kind = neumann_type, c.n = n_faces;
    for (int d = 0; d < 3; ++d) c.onoff[d] = onoff[d], c.val[d] = val[d];
    if (to_device(ctx, &c.d_faces, face_nodes, n_faces * ctx->face_nn)) return B200S_ERR_CUDA;
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->neumann.push_back(c);
    return (int)ctx->neumann.size() - 1;
  }

  int b200s_add_volume_neumann(b200s_ctx* ctx, int64_t n_eles, const int32_t* ele_ids, const int32_t* onoff,
      const double* val)
  {
    if (!ctx || !onoff || !val || n_eles < 0) return B200S_ERR_INVALID;
    if (ctx->celltype < 0) return fail(ctx, B200S_ERR_INVALID, "set_mesh must precede add_volume_neumann");
    cudaSetDevice(ctx->device);
    b200s_ctx::NeumannCond c;
    c.volume = true, c.kind = B200S_NEUMANN_LIVE, c.n = ele_ids ? n_eles : ctx->n_col_eles;
    for (int d = 0; d < 3; ++d) c.onoff[d] = onoff[d], c.val[d] = val[d];
    if (ele_ids)
    {
      std::vector<int32_t> slot((size_t)ctx->n_col_eles, -1);
      for (int64_t k = 0; k < n_eles; ++k)
      {
        if (ele_ids[k] < 0 || ele_ids[k] >= ctx->n_col_eles)
          return fail(ctx, B200S_ERR_INVALID, "element id %d out of range", ele_ids[k]);
        slot[ele_ids[k]] = (int32_t)k;
      }
      if (to_device(ctx, &c.d_ele_slot, slot.data(), ctx->n_col_eles)) return B200S_ERR_CUDA;
      CK(cudaStreamSynchronize(ctx->stream));
    }
    ctx->neumann.push_back(c);
    return (int)ctx->neumann.size() - 1;
  }

  int64_t b200s_neumann_gp_count(const b200s_ctx* ctx, int cond)
  {
    if (!ctx || cond < 0 || cond >= (int)ctx->neumann.size()) return -1;
    const auto& c = ctx->neumann[cond];
    return c.n * (c.volume ? ctx->ngp : ctx->face_ngp);
  }

  int b200s_set_neumann_factors(b200s_ctx* ctx, int cond, const double* dof_fac, const double* gp_fac)
  {
    if (!ctx || cond < 0 || cond >= (int)ctx->neumann.size()) return B200S_ERR_INVALID;
    cudaSetDevice(ctx->device);
    auto& c = ctx->neumann[cond];
    for (int d = 0; d < 3; ++d) c.fac[d] = dof_fac ? dof_fac[d] : 1.0;
    if (!gp_fac)
    {
      CK(cudaStreamSynchronize(ctx->stream));
      dfree(c.d_gp_fac);
      return B200S_OK;
    }
    const int64_t n = 3 * b200s_neumann_gp_count(ctx, cond);
    if (!c.d_gp_fac && to_device(ctx, &c.d_gp_fac, (const double*)nullptr, n)) return B200S_ERR_CUDA;
    CK(cudaMemcpyAsync(c.d_gp_fac, gp_fac, sizeof(double) * (size_t)n, cudaMemcpyDefault, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return B200S_OK;
  }

  // (static helpers of the entry points below)
    static NeumannSurfArgs neumann_surface_args(b200s_ctx* ctx, const b200s_ctx::NeumannCond& c, const double* d_u_col,
        int flags)
    {
      NeumannSurfArgs A;
      A.T.nn = ctx->face_nn, A.T.ngp = ctx->face_ngp;
      A.T.N = ctx->d_face_N, A.T.d0 = ctx->d_face_d0, A.T.d1 = ctx->d_face_d1, A.T.w = ctx->d_face_w;
      A.X = ctx->dX;
      // configuration (:127-178): material for Live and for an orthopressure while MULF is active
      const bool material = c.kind == B200S_NEUMANN_LIVE ||
                            (c.kind == B200S_NEUMANN_ORTHOPRESSURE && (flags & B200S_NEUMANN_MULF_ACTIVE));
      A.u = material ? nullptr : d_u_col;
      A.faces = c.d_faces, A.n_faces = c.n, A.kind = c.kind;
      for (int d = 0; d < 3; ++d) A.onoff[d] = c.onoff[d], A.val[d] = c.val[d] * c.fac[d];
      A.gp_fac = c.d_gp_fac;
      A.node_row = ctx->d_node_row, A.node_colbase = ctx->d_node_colbase, A.colind = ctx->d_colind;
      A.fext = ctx->d_fext;
      const bool lin = c.kind == B200S_NEUMANN_ORTHOPRESSURE && (flags & B200S_NEUMANN_LOADLIN) && !material;
      A.K = lin ? ctx->d_vals : nullptr;
      return A;
    }
    static NeumannVolArgs neumann_volume_args(b200s_ctx* ctx, const b200s_ctx::NeumannCond& c)
    {
      NeumannVolArgs A;
      A.X = ctx->dX, A.conn = ctx->dconn, A.ele_id = ctx->d_ele_id, A.ele_slot = c.d_ele_slot;
      A.N_gp = ctx->d_N_gp, A.dN_gp = ctx->d_dN_gp, A.w = ctx->d_w;
      A.n_ele = ctx->n_col_eles, A.nn = ctx->nn, A.ngp = ctx->ngp;
      for (int d = 0; d < 3; ++d) A.onoff[d] = c.onoff[d], A.val[d] = c.val[d] * c.fac[d];
      A.gp_fac = c.d_gp_fac;
      A.node_row = ctx->d_node_row, A.fext = ctx->d_fext;
      return A;
    }
    // column-map displacement the deformation dependent conditions act on
    static int neumann_state(b200s_ctx* ctx, const double* u_row, const double** d_u_col)
    {
      if (!u_row)
      {
        const int rc = import_row_to_col(ctx, ctx->d_u_row, ctx->d_u_col, false);
        *d_u_col = ctx->d_u_col;
        return rc;
      }
      if (stage_row_vector(ctx, &ctx->d_nm_row, u_row)) return B200S_ERR_CUDA;
      if (!ctx->d_nm_col)
      {
        if (to_device(ctx, &ctx->d_nm_col, (const double*)nullptr, 3 * ctx->n_col_nodes)) return B200S_ERR_CUDA;
        CK(cudaMemsetAsync(ctx->d_nm_col, 0, sizeof(double) * 3 * (size_t)ctx->n_col_nodes, ctx->stream));
      }
      *d_u_col = ctx->d_nm_col;
      return import_row_to_col(ctx, ctx->d_nm_row, ctx->d_nm_col, false);
    }

  int b200s_neumann_gp_coords(b200s_ctx* ctx, int cond, int flags, double* xyz_out)
  {
    if (!ctx || !xyz_out || cond < 0 || cond >= (int)ctx->neumann.size()) return B200S_ERR_INVALID;
    if (!ctx->d_rowptr) return fail(ctx, B200S_ERR_INVALID, "set_graph must precede neumann_gp_coords");
    cudaSetDevice(ctx->device);
    const auto& c = ctx->neumann[cond];
    const int64_t n = b200s_neumann_gp_count(ctx, cond);
    double* d_out = nullptr;
    if (to_device(ctx, &d_out, (const double*)nullptr, 3 * n)) return B200S_ERR_CUDA;
    int rc = B200S_OK;
    if (c.volume)
    {
      NeumannVolArgs A = neumann_volume_args(ctx, c);
      if (n) k_neumann_volume_gp_coords<<<nblk(ctx->n_col_eles * ctx->ngp), 256, 0, ctx->stream>>>(A, d_out);
    }
    else
    {
      const double* d_u_col = nullptr;
      if (c.kind != B200S_NEUMANN_LIVE) rc = neumann_state(ctx, nullptr, &d_u_col);
      NeumannSurfArgs A = neumann_surface_args(ctx, c, d_u_col, flags);
      if (!rc && n) k_neumann_surface_gp_coords<<<nblk(n), 256, 0, ctx->stream>>>(A, d_out);
    }
    ctx->launches += 1;
    cudaError_t e = cudaMemcpyAsync(xyz_out, d_out, sizeof(double) * 3 * (size_t)n, cudaMemcpyDefault, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e == cudaSuccess) e = cudaGetLastError();
    cudaFree(d_out);
    if (rc) return rc;
    if (e != cudaSuccess) return fail(ctx, B200S_ERR_CUDA, "neumann_gp_coords: %s", cudaGetErrorString(e));
    return B200S_OK;
  }

  int b200s_evaluate_neumann(b200s_ctx* ctx, const double* u_row, int flags, double* fext_row)
  {
    if (!ctx) return B200S_ERR_INVALID;
    if (!ctx->d_rowptr || !ctx->d_node_row)
      return fail(ctx, B200S_ERR_INVALID, "set_mesh and set_graph must precede evaluate_neumann");
    cudaSetDevice(ctx->device);
    cudaStream_t st = ctx->stream;
    if (!ctx->d_fext && to_device(ctx, &ctx->d_fext, (const double*)nullptr, ctx->n_rows)) return B200S_ERR_CUDA;
    CK(cudaMemsetAsync(ctx->d_fext, 0, sizeof(double) * (size_t)ctx->n_rows, st));
    bool needs_state = false, needs_matrix = false;
    for (const auto& c : ctx->neumann)
    {
      if (!c.volume && c.kind != B200S_NEUMANN_LIVE) needs_state = true;
      if (!c.volume && c.kind == B200S_NEUMANN_ORTHOPRESSURE && (flags & B200S_NEUMANN_LOADLIN)) needs_matrix = true;
    }
    if (needs_matrix && !ctx->d_vals)
      return fail(ctx, B200S_ERR_INVALID, "load linearisation needs the stiffness matrix of this iteration");
    const double* d_u_col = nullptr;
    if (needs_state)
    {
      const int rc = neumann_state(ctx, u_row, &d_u_col);
      if (rc) return rc;
    }
    int* d_flag = nullptr;
    if (needs_matrix)
    {
      if (!ctx->d_nm_flag && to_device(ctx, &ctx->d_nm_flag, (const int*)nullptr, 1)) return B200S_ERR_CUDA;
      d_flag = ctx->d_nm_flag;
      CK(cudaMemsetAsync(d_flag, 0, sizeof(int), st));
    }
    for (const auto& c : ctx->neumann)
    {
      if (c.n == 0) continue;
      if (c.volume)
      {
        NeumannVolArgs A = neumann_volume_args(ctx, c);
        k_neumann_volume_force<<<nblk(ctx->n_col_eles * ctx->nn), 256, 0, st>>>(A);
        ctx->launches += 1;
        continue;
      }
      NeumannSurfArgs A = neumann_surface_args(ctx, c, d_u_col, flags);
      k_neumann_surface_force<<<nblk(c.n * ctx->face_nn), 256, 0, st>>>(A);
      ctx->launches += 1;
      if (A.K)
      {
        k_neumann_surface_lin<<<nblk(c.n * ctx->face_nn * ctx->face_nn), 256, 0, st>>>(A, d_flag);
        ctx->launches += 1;
      }
    }
    int h_flag = 0;
    cudaError_t e = cudaSuccess;
    if (d_flag) e = cudaMemcpyAsync(&h_flag, d_flag, sizeof(int), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess && fext_row && fext_row != ctx->d_fext)
      e = cudaMemcpyAsync(fext_row, ctx->d_fext, sizeof(double) * (size_t)ctx->n_rows, cudaMemcpyDefault, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) return fail(ctx, B200S_ERR_CUDA, "evaluate_neumann: %s", cudaGetErrorString(e));
    if (h_flag)
      return fail(ctx, B200S_ERR_INVALID,
          "a loaded face couples two nodes without an entry in the stiffness graph (face not on an element of the block?)");
    return B200S_OK;
  }
  double* b200s_device_fext(b200s_ctx* ctx) { return ctx ? ctx->d_fext : nullptr; }


  int b200s_apply_dirichlet(b200s_ctx* ctx, int which, const double* dbctoggle_row, int diagonalblock)
  {
    if (!ctx || !dbctoggle_row) return B200S_ERR_INVALID;
    if (!ctx->d_rowptr || !ctx->d_colind || !matrix_values(ctx, which))
      return fail(ctx, B200S_ERR_INVALID, "apply_dirichlet needs an assembled matrix");
    cudaSetDevice(ctx->device);
    if (stage_row_vector(ctx, &ctx->d_toggle, dbctoggle_row)) return B200S_ERR_CUDA;
    if (!ctx->d_nm_flag && to_device(ctx, &ctx->d_nm_flag, (const int*)nullptr, 1)) return B200S_ERR_CUDA;
    CK(cudaMemsetAsync(ctx->d_nm_flag, 0, sizeof(int), ctx->stream));
    k_apply_dirichlet<<<nblk(32 * ctx->n_rows), 256, 0, ctx->stream>>>(ctx->n_rows, ctx->d_rowptr,
        ctx->d_colind, ctx->d_row_diagcol, ctx->d_toggle, diagonalblock, matrix_values(ctx, which), ctx->d_nm_flag);
    ctx->launches += 1;
    int missing = 0;
    CK(cudaMemcpyAsync(&missing, ctx->d_nm_flag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaGetLastError());
    if (missing)
      return fail(ctx, B200S_ERR_INVALID, "apply_dirichlet: %d Dirichlet rows have no diagonal entry in the graph",
          missing);
    return B200S_OK;
  }

  int b200s_apply_dirichlet_to_system(b200s_ctx* ctx, double* x_row, double* b_row,
      const double* dbcval_row, const double* dbctoggle_row)
  {
    if (!ctx || !dbcval_row || !dbctoggle_row || (!x_row && !b_row)) return B200S_ERR_INVALID;
    if (!ctx->d_rowptr) return fail(ctx, B200S_ERR_INVALID, "set_graph first");
    cudaSetDevice(ctx->device);
    cudaStream_t st = ctx->stream;
    const size_t bytes = sizeof(double) * (size_t)ctx->n_rows;
    // toggle in d_toggle, values in d_la_y, the vectors pass through d_la_x one after the other
    if (stage_row_vector(ctx, &ctx->d_toggle, dbctoggle_row)) return B200S_ERR_CUDA;
    if (stage_row_vector(ctx, &ctx->d_la_y, dbcval_row)) return B200S_ERR_CUDA;
    for (double* v : {x_row, b_row})
    {
      if (!v) continue;
      if (stage_row_vector(ctx, &ctx->d_la_x, v)) return B200S_ERR_CUDA;
      k_dirichlet_to_system<<<nblk(ctx->n_rows), 256, 0, st>>>(ctx->n_rows, ctx->d_toggle, ctx->d_la_y,
          ctx->d_la_x, nullptr);
      ctx->launches += 1;
      CK(cudaMemcpyAsync(v, ctx->d_la_x, bytes, cudaMemcpyDefault, st));
    }
    CK(cudaStreamSynchronize(st));
    CK(cudaGetLastError());
    return B200S_OK;
  }

  // y = A x on device vectors (x on the row map = domain map); dot/w as in the kernels
  static int multiply_device(b200s_ctx* ctx, int which, const double* d_x_row, double* d_y_row,
      const double* d_w, double* d_dot)
  {
    const double* vals = matrix_values(ctx, which);
    if (!ctx->d_rowptr || !ctx->d_colind || !vals)
      return fail(ctx, B200S_ERR_INVALID, "multiply needs an assembled matrix");
    if (!ctx->d_xmat && to_device(ctx, &ctx->d_xmat, (const double*)nullptr, ctx->n_mat_cols))
      return B200S_ERR_CUDA;
    const int rc = import_row_to_col(ctx, d_x_row, ctx->d_xmat, true);
    if (rc) return rc;
    if (ctx->node3_ok)
      k_spmv_node3<<<nblk(32 * ctx->n_col_nodes), 256, 0, ctx->stream>>>(ctx->n_col_nodes, ctx->d_node_row,
          ctx->d_blkcol, vals, ctx->d_xmat, d_y_row, d_w, d_dot);
    else
      k_spmv_csr<<<nblk(32 * ctx->n_rows), 256, 0, ctx->stream>>>(ctx->n_rows, ctx->d_rowptr, ctx->d_colind,
          vals, ctx->d_xmat, d_y_row, d_w, d_dot);
    ctx->launches += 1;
    return B200S_OK;
  }

  int b200s_multiply(b200s_ctx* ctx, int which, const double* x_row, double* y_row)
  {
    if (!ctx || !x_row || !y_row) return B200S_ERR_INVALID;
    if (!ctx->d_rowptr) return fail(ctx, B200S_ERR_INVALID, "set_graph first");
    cudaSetDevice(ctx->device);
    cudaStream_t st = ctx->stream;
    if (stage_row_vector(ctx, &ctx->d_la_x, x_row)) return B200S_ERR_CUDA;
    if (!ctx->d_la_y && to_device(ctx, &ctx->d_la_y, (const double*)nullptr, ctx->n_rows)) return B200S_ERR_CUDA;
    CK(cudaEventRecord(ctx->ev_la[0], st));
    const int rc = multiply_device(ctx, which, ctx->d_la_x, ctx->d_la_y, nullptr, nullptr);
    if (rc) return rc;
    CK(cudaEventRecord(ctx->ev_la[1], st));
    CK(cudaMemcpyAsync(y_row, ctx->d_la_y, sizeof(double) * (size_t)ctx->n_rows, cudaMemcpyDefault, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaGetLastError());
    float ms = 0;
    cudaEventElapsedTime(&ms, ctx->ev_la[0], ctx->ev_la[1]);
    ctx->last_spmv_ms = ms;
    return B200S_OK;
  }
  double b200s_last_multiply_ms(const b200s_ctx* ctx) { return ctx ? ctx->last_spmv_ms : 0.0; }

  int b200s_solve_cg(b200s_ctx* ctx, const double* b_row, double* x_row, double rtol, int maxit,
      int* iters_out, double* relres_out)
  {
    if (!ctx || !b_row || !x_row || maxit < 1) return B200S_ERR_INVALID;
    if (!ctx->d_rowptr || !ctx->d_vals) return fail(ctx, B200S_ERR_INVALID, "no assembled matrix");
    cudaSetDevice(ctx->device);
    cudaStream_t st = ctx->stream;
    const int64_t n = ctx->n_rows, nrn = n / 3;
    const size_t bytes = sizeof(double) * (size_t)n;
    for (int i = 0; i < 4; ++i)
      if (!ctx->d_cg[i] && to_device(ctx, &ctx->d_cg[i], (const double*)nullptr, n)) return B200S_ERR_CUDA;
    if (!ctx->d_cg[4] && to_device(ctx, &ctx->d_cg[4], (const double*)nullptr, 3 * n)) return B200S_ERR_CUDA;
    if (!ctx->d_cg_s && to_device(ctx, &ctx->d_cg_s, (const double*)nullptr, 4)) return B200S_ERR_CUDA;
    if (ctx->cg_hist_len < maxit + 1)
    {
      dfree(ctx->d_cg_hist);
      if (to_device(ctx, &ctx->d_cg_hist, (const double*)nullptr, maxit + 1)) return B200S_ERR_CUDA;
      ctx->cg_hist_len = maxit + 1;
    }
    double *r = ctx->d_cg[0], *z = ctx->d_cg[1], *p = ctx->d_cg[2], *Ap = ctx->d_cg[3], *dinv = ctx->d_cg[4];
    double* s = ctx->d_cg_s;
    // x0 = caller's x (initial guess), b staged in d_la_y, x in d_la_x
    if (stage_row_vector(ctx, &ctx->d_la_x, x_row)) return B200S_ERR_CUDA;
    if (stage_row_vector(ctx, &ctx->d_la_y, b_row)) return B200S_ERR_CUDA;
    double* x = ctx->d_la_x;
    const double* b = ctx->d_la_y;
    auto allreduce = [&](double* v, int cnt) -> int
    {
      if (ctx->nranks <= 1 || !ctx->comm) return 0;
      const int rc = ctx->nccl.AllReduce(v, v, (size_t)cnt, /*ncclDouble*/ 8, /*ncclSum*/ 0, ctx->comm, st);
      return rc ? fail(ctx, B200S_ERR_COMM, "allreduce: %s", ctx->nccl.GetErrorString(rc)) : 0;
    };
    k_block_diag_inverse<<<nblk(ctx->n_col_nodes), 256, 0, st>>>(
        ctx->n_col_nodes, ctx->d_node_row, ctx->d_node_colbase, ctx->d_colind, ctx->d_vals, dinv);
    CK(cudaMemsetAsync(s, 0, 4 * sizeof(double), st));
    int rc = multiply_device(ctx, B200S_MATRIX_STIFFNESS, x, Ap, nullptr, nullptr);
    if (rc) return rc;
    k_residual<<<nblk(n), 256, 0, st>>>(n, b, Ap, r);
    k_cg_update<<<nblk(nrn), 256, 0, st>>>(nrn, dinv, p, Ap, x, r, z, s, 1);
    if (allreduce(s + 2, 2)) return B200S_ERR_COMM;
    k_cg_direction<<<nblk(n), 256, 0, st>>>(n, z, p, s, 1);
    k_cg_roll<<<1, 1, 0, st>>>(s, ctx->d_cg_hist, 0);
    ctx->launches += 5;
    double rr0 = 0.0, rr = 0.0;
    CK(cudaMemcpyAsync(&rr0, ctx->d_cg_hist, sizeof(double), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    rr = rr0;
    int it = 0;
    const int check = 8;  // convergence is read back every `check` iterations (no sync in between)
    while (it < maxit && rr0 > 0.0 && std::sqrt(rr / rr0) > rtol)
    {
      const int stop = it + check < maxit ? it + check : maxit;
      for (; it < stop; ++it)
      {
        rc = multiply_device(ctx, B200S_MATRIX_STIFFNESS, p, Ap, p, s);
        if (rc) return rc;
        if (allreduce(s, 1)) return B200S_ERR_COMM;
        k_cg_update<<<nblk(nrn), 256, 0, st>>>(nrn, dinv, p, Ap, x, r, z, s, 0);
        if (allreduce(s + 2, 2)) return B200S_ERR_COMM;
        k_cg_direction<<<nblk(n), 256, 0, st>>>(n, z, p, s, 0);
        k_cg_roll<<<1, 1, 0, st>>>(s, ctx->d_cg_hist, it + 1);
        ctx->launches += 3;
      }
      CK(cudaMemcpyAsync(&rr, ctx->d_cg_hist + it, sizeof(double), cudaMemcpyDeviceToHost, st));
      CK(cudaStreamSynchronize(st));
      if (!(rr == rr)) break;  // NaN: breakdown (indefinite matrix)
    }
    CK(cudaMemcpyAsync(x_row, x, bytes, cudaMemcpyDefault, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaGetLastError());
    if (iters_out) *iters_out = it;
    const double rel = rr0 > 0.0 ? std::sqrt(rr / rr0) : 0.0;
    if (relres_out) *relres_out = rel;
    if (!(rel == rel)) return fail(ctx, B200S_ERR_NONFINITE, "CG breakdown after %d iterations", it);
    return B200S_OK;
  }

  int b200s_measure_fp64_peak(int device, double* tflops, double* sm_mhz)
  {
    if (cudaSetDevice(device) != cudaSuccess) return B200S_ERR_CUDA;
    int n_sm = 0;
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, device);
    double* d = nullptr;
    if (cudaMalloc((void**)&d, 8) != cudaSuccess) return B200S_ERR_CUDA;
    cudaEvent_t a, b;
    cudaEventCreate(&a), cudaEventCreate(&b);
    const int iters = 4096, blocks = n_sm * 8, threads = 256;
    double best = 0;
    for (int rep = 0; rep < 6; ++rep)
    {
      cudaEventRecord(a);
      k_fp64_peak<<<blocks, threads>>>(d, iters, 0.999999, 1e-7);
      cudaEventRecord(b);
      cudaEventSynchronize(b);
      float ms = 0;
      cudaEventElapsedTime(&ms, a, b);
      const double fl = 2.0 * 64.0 * iters * (double)blocks * threads;
      const double tf = fl / (ms * 1e-3) / 1e12;
      if (rep > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(a), cudaEventDestroy(b), cudaFree(d);
    if (cudaGetLastError() != cudaSuccess) return B200S_ERR_CUDA;
    if (tflops) *tflops = best;
    if (sm_mhz) *sm_mhz = best * 1e12 / (2.0 * 64.0 * n_sm) / 1e6;  // clock implied by 64 DFMA/clk/SM
    return B200S_OK;
  }

  // Diagnostics used by tools/ and DESIGN.md measurements (not part of the hot path).
  //   which 0: DFMA TFLOP/s, 1: DMMA m8n8k4 TFLOP/s, 2: DMMA+DFMA interleaved TFLOP/s,
  //   3: cudaMemset GB/s over `size` bytes, 4/5/6: red.add.f64 G atomics/s for pattern 0/1/2
  int b200s_microbench(int device, int which, int64_t size, double* result)
  {
    if (!result || cudaSetDevice(device) != cudaSuccess) return B200S_ERR_CUDA;
    int n_sm = 0;
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, device);
    cudaEvent_t a, b;
    cudaEventCreate(&a), cudaEventCreate(&b);
    double best = 0;
    float ms = 0;
    if (which == 0) return b200s_measure_fp64_peak(device, result, nullptr);
    if (which == 1 || which == 2)
    {
      double* d = nullptr;
      cudaMalloc((void**)&d, 8);
      const int iters = 2048, blocks = n_sm * 8, threads = 256;
      for (int rep = 0; rep < 5; ++rep)
      {
        cudaEventRecord(a);
        if (which == 1)
          k_dmma_peak<1><<<blocks, threads>>>(d, iters, 0.999999, 1e-7);
        else
          k_dmma_peak<2><<<blocks, threads>>>(d, iters, 0.999999, 1e-7);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        cudaEventElapsedTime(&ms, a, b);
        // per warp and inner step: 4 DMMA x 256 FMA (+ 8 DFMA x 32 lanes)
        const double fma_per_warp = iters * 4.0 * (4 * 256.0 + (which == 2 ? 8 * 32.0 : 0.0));
        const double tf = 2.0 * fma_per_warp * blocks * (threads / 32) / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
      }
      cudaFree(d);
    }
    else if (which == 3)
    {
      void* d = nullptr;
      if (cudaMalloc(&d, (size_t)size) != cudaSuccess) return B200S_ERR_CUDA;
      for (int rep = 0; rep < 5; ++rep)
      {
        cudaEventRecord(a);
        cudaMemsetAsync(d, 0, (size_t)size);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        cudaEventElapsedTime(&ms, a, b);
        const double gbs = size / (ms * 1e-3) / 1e9;
        if (rep > 0 && gbs > best) best = gbs;
      }
      cudaFree(d);
    }
    else if ((which >= 4 && which <= 6) || which >= 100)
    {
      double* d = nullptr;
      const int64_t n = size / 8;
      if (cudaMalloc((void**)&d, (size_t)size) != cudaSuccess) return B200S_ERR_CUDA;
      cudaMemset(d, 0, (size_t)size);
      const int blocks = n_sm * 16, threads = 256, per_thread = 96;
      for (int rep = 0; rep < 4; ++rep)
      {
        cudaEventRecord(a);
        if (which >= 100)
        {
          unsigned np2 = 1;
          while ((int64_t)np2 * 2 <= n - (1 << 22) - 64) np2 *= 2;
          switch (which - 100)
          {
#define RUNCASE(R) case R: k_red_runs<R><<<blocks, threads>>>(d, np2, per_thread); break
            RUNCASE(1); RUNCASE(2); RUNCASE(3); RUNCASE(4); RUNCASE(6); RUNCASE(8); RUNCASE(12);
            RUNCASE(16); RUNCASE(24); RUNCASE(32);
#undef RUNCASE
            default: cudaFree(d); return B200S_ERR_INVALID;
          }
        }
        else
          k_red_pattern<<<blocks, threads>>>(d, n, which - 4, per_thread);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        cudaEventElapsedTime(&ms, a, b);
        const double g = (double)blocks * threads * per_thread / (ms * 1e-3) / 1e9;
        if (rep > 0 && g > best) best = g;
      }
      cudaFree(d);
    }
    else
      return B200S_ERR_INVALID;
    cudaEventDestroy(a), cudaEventDestroy(b);
    if (cudaGetLastError() != cudaSuccess) return B200S_ERR_CUDA;
    *result = best;
    return B200S_OK;
  }

  // ---- pinned host staging ----------------------------------------------------------------------------
  // NUMA node of a CUDA device (sysfs numa_node of its PCI function), -1 if unknown / single node
  static int device_numa_node(int device)
  {
    char bus[32] = {0};
    if (cudaDeviceGetPCIBusId(bus, sizeof bus, device) != cudaSuccess) return -1;
    for (char* c = bus; *c; ++c) *c = (char)tolower(*c);
    char path[128];
    snprintf(path, sizeof path, "/sys/bus/pci/devices/%s/numa_node", bus);
    FILE* f = fopen(path, "r");
    if (!f) return -1;
    int node = -1;
    if (fscanf(f, "%d", &node) != 1) node = -1;
    fclose(f);
    return node;
  }
  struct HostBlock
  {
    void* p;
    size_t bytes;
    bool mapped;  // mmap + cudaHostRegister (NUMA placed) instead of cudaMallocHost
  };
  static std::vector<HostBlock>& host_blocks()
  {
    static std::vector<HostBlock> v;
    return v;
  }
  static std::mutex& host_blocks_mutex()  // contexts of several ranks may live in one process (threads)
  {
    static std::mutex m;
    return m;
  }

  void* b200s_host_alloc(int64_t bytes)
  {
    void* p = nullptr;
    if (bytes <= 0 || cudaMallocHost(&p, (size_t)bytes) != cudaSuccess) return nullptr;
    std::lock_guard<std::mutex> lk(host_blocks_mutex());
    host_blocks().push_back({p, (size_t)bytes, false});
    return p;
  }

  // Pinned staging memory placed on the NUMA node the GPU hangs off: with one rank per GPU every rank DMAs its
  // K values (GBs per Newton iteration) into host memory, and pages that all sit on one socket serialise the
  // ranks on that socket's memory controllers and the inter-socket link.  mmap + mbind(MPOL_BIND) + first touch
  // + cudaHostRegister; falls back to cudaMallocHost when the node is unknown or binding is not permitted.
  void* b200s_host_alloc_near(int device, int64_t bytes, int* numa_node_out)
  {
    if (numa_node_out) *numa_node_out = -1;
    if (bytes <= 0) return nullptr;
    const int node = device_numa_node(device);
    if (node >= 0 && node < 1024)
    {
      const size_t len = ((size_t)bytes + (2u << 20) - 1) & ~(size_t)((2u << 20) - 1);
      void* p = mmap(nullptr, len, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
      if (p != MAP_FAILED)
      {
        unsigned long mask[16] = {0};
        mask[node / (8 * sizeof(unsigned long))] |= 1ul << (node % (8 * sizeof(unsigned long)));
        const long rc = syscall(SYS_mbind, p, len, /*MPOL_BIND*/ 2, mask, (unsigned long)(8 * sizeof mask), 0ul);
        if (rc == 0)
        {
          madvise(p, len, MADV_HUGEPAGE);
          for (size_t o = 0; o < len; o += 4096) ((volatile char*)p)[o] = 0;  // first touch on the bound node
          cudaSetDevice(device);
          if (cudaHostRegister(p, len, cudaHostRegisterPortable) == cudaSuccess)
          {
            std::lock_guard<std::mutex> lk(host_blocks_mutex());
            host_blocks().push_back({p, len, true});
            if (numa_node_out) *numa_node_out = node;
            return p;
          }
          cudaGetLastError();
        }
        munmap(p, len);
      }
    }
    return b200s_host_alloc(bytes);
  }

  void b200s_host_free(void* p)
  {
    if (!p) return;
    HostBlock blk{p, 0, false};
    {
      std::lock_guard<std::mutex> lk(host_blocks_mutex());
      auto& v = host_blocks();
      for (size_t i = 0; i < v.size(); ++i)
        if (v[i].p == p)
        {
          blk = v[i];
          v.erase(v.begin() + (long)i);
          break;
        }
    }
    if (blk.mapped)
    {
      cudaHostUnregister(p);
      munmap(p, blk.bytes);
    }
    else
      cudaFreeHost(p);
  }

  // diagnostics: device element order (caller's column-element index of every device element)
  int b200s_debug_element_order(b200s_ctx* ctx, int32_t* out)
  {
    if (!ctx || !out) return B200S_ERR_INVALID;
    cudaSetDevice(ctx->device);
    if (!ctx->d_ele_id)
    {
      for (int64_t e = 0; e < ctx->n_col_eles; ++e) out[e] = (int32_t)e;
      return B200S_OK;
    }
    CK(cudaMemcpy(out, ctx->d_ele_id, 4 * (size_t)ctx->n_col_eles, cudaMemcpyDeviceToHost));
    return B200S_OK;
  }
  int64_t b200s_last_d2h_value_bytes(const b200s_ctx* ctx) { return ctx ? ctx->last_d2h_bytes : 0; }
  const char* b200s_symmetric_transfer_status(const b200s_ctx* ctx)
  {
    if (!ctx || !ctx->sym) return "not prepared";
    return ctx->sym->usable ? "in use" : ctx->sym->why.c_str();
  }
  // what the last evaluate did: out[0] element-kernel launches, [1] value chunks copied to the host while later
  // segments were still being evaluated, [2] 1 if the ghost import ran beside the interior elements,
  // [3] row phases of the installed graph, [4] elements holding a ghost node
  int b200s_last_evaluate_info(const b200s_ctx* ctx, int64_t* out5)
  {
    if (!ctx || !out5) return B200S_ERR_INVALID;
    out5[0] = ctx->last_kernel_launches, out5[1] = ctx->last_pipeline_chunks, out5[2] = ctx->last_halo_overlapped;
    out5[3] = ctx->n_phases, out5[4] = ctx->segmented && ctx->seg_begin.size() > 1 ? ctx->seg_begin[1] : 0;
    return B200S_OK;
  }
}
#endif  // B200S_PART_CORE
