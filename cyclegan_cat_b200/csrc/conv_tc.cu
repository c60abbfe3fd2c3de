// tcgen05 / TMEM / TMA implicit-GEMM convolution kernels for sm_100a (bf16 in, fp32 accumulate).
//
//   conv_tc_kernel   forward conv and data-gradient (a conv with the transposed weights):
//                    D[128 pixels x BN channels] += sum over (tap, 64-channel chunk) A_tap[128x64] * B_tap[BNx64]^T
//                    A tiles are plain TMA box loads of the NHWC activation (one box per filter tap, shifted
//                    coordinates; zero padding comes from TMA out-of-bounds fill), B tiles are TMA loads of the
//                    packed bf16 weights.  Both operands K-major, SWIZZLE_128B.  Persistent CTAs, 4-5 stage
//                    smem ring, double-buffered TMEM accumulator, warp-specialised:
//                    warp 0 = TMA producer, warp 1 = MMA issuer (one elected thread), warps 2-5 = epilogue.
//   wgrad_tc_kernel  weight gradient dW[tap][ci][co] = sum_pixels X[pixel+tap][ci] * dY[pixel][co]:
//                    the SAME activation tiles, now used as MN-major operands (K = pixels), split-K over the
//                    pixel range, fp32 vector reductions (red.global.add.v4.f32) into the float32 gradient.
#include <cuda.h>
#include <stdlib.h>

#include "conv_tc.h"
#include "ptx_async.h"
#include "prof.h"

// ------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                                            int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                                            int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem], bf16 x bf16 -> fp32, issued by ONE thread for the whole CTA
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// One lane of a CONVERGED warp (elect.sync).  The MMA-issuing code is written as `if (elect_one()) { tcgen05.mma ... }` inside
// warp-uniform control flow: operands that are uniform by data flow then stay in uniform registers.  Under `if (lane == 0)`
// the compiler cannot prove uniformity and wraps every tcgen05 instruction in an ELECT / BRA.U.ANY loop (~25 instructions per
// MMA; ncu showed the MMA warp of the narrow window layers executing instructions 100 % of the time, ~3000 cycles per tile).
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred P;\n\t"
        "elect.sync _|P, 0xffffffff;\n\t"
        "selp.b32 %0, 1, 0, P;\n\t}"
        : "=r"(pred));
    return pred != 0;
}
// arrive on an mbarrier when all MMAs issued so far by this thread have completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// 32 lanes x 32 consecutive fp32 columns -> 32 registers per thread (thread = lane, register = column)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): SWIZZLE_128B, version 1 (sm_100)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);              // start address, bits [0,14)
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;     // leading byte offset, bits [16,30)
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;     // stride byte offset, bits [32,46)
    d |= (uint64_t)1 << 46;                               // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                               // layout type: SWIZZLE_128B
    return d;
}
// same for SWIZZLE_32B K-major tiles: rows of 32 bytes (16 bf16 = one MMA K), 8-row atoms of 256 bytes
__device__ __forceinline__ uint64_t make_smem_desc32(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;                               // LBO (unused for swizzled K-major)
    d |= (uint64_t)(256 >> 4) << 32;                      // SBO = 256 B
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)6 << 61;                               // layout type: SWIZZLE_32B
    return d;
}
// instruction descriptor (cute::UMMA::InstrDescriptor): BF16 x BF16 -> F32, dense
static inline uint32_t make_idesc(int M, int N, int a_mn_major, int b_mn_major) {
    uint32_t d = 0;
    d |= 1u << 4;                       // c_format = F32
    d |= 1u << 7;                       // a_format = BF16
    d |= 1u << 10;                      // b_format = BF16
    d |= (uint32_t)a_mn_major << 15;
    d |= (uint32_t)b_mn_major << 16;
    d |= (uint32_t)(N >> 3) << 17;
    d |= (uint32_t)(M >> 4) << 24;
    return d;
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&t);
}

// column sums of a 32 x 32 tile held one row per lane (v[j] = column j of this lane's row): recursive halving with
// shuffles -- after the 5 steps lane L holds the sum of column L.  31 shuffles.
// (a.lo + b.lo, a.hi + b.hi) of two packed bf16 pairs, summed in fp32 and rounded once
__device__ __forceinline__ uint32_t add_bf16x2(uint32_t a, uint32_t b) {
    const float lo = __uint_as_float(a << 16) + __uint_as_float(b << 16);
    const float hi = __uint_as_float(a & 0xffff0000u) + __uint_as_float(b & 0xffff0000u);
    return pack_bf16x2(lo, hi);
}

__device__ __forceinline__ float warp_colsum32(float (&s)[32], int lane) {
#pragma unroll
    for (int off = 16, n = 32; off >= 1; off >>= 1, n >>= 1) {
        const bool up = (lane & off) != 0;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            if (j < n / 2) {
                const float send = up ? s[j] : s[j + n / 2];
                const float keep = up ? s[j + n / 2] : s[j];
                s[j] = keep + __shfl_xor_sync(0xffffffffu, send, off);
            }
        }
    }
    return s[0];
}

// Window mode: zero the 16-byte granules of a freshly landed K-major SWIZZLE_128B tile ([box_h][box_w] rows of 128 bytes,
// row = pixel) whose window pixel lies outside the image row.  Row (hb, wb) holds elements [c0, c0+64) of the k*C-element
// window that starts `pl` pixels left of pixel p = w0 + wb; window pixel q = e / C is column p - pl + q.  Only rows within
// pl of the left edge or k-1-pl of the right edge have anything to zero, so a 128-wide tile of a 256-wide row touches <= 3
// rows per step.  Called by all 32 lanes of the MMA warp between the TMA-complete wait and the MMA issue; the
// fence.proxy.async orders the generic-proxy stores before the tensor core's async-proxy reads.
__device__ __forceinline__ void win_fix(uint32_t tile, int box_w, int box_h, int w0, int W, int C, int k, int pl, int c0,
                                        int lane) {
    const int nl = w0 < pl ? min(pl - w0, box_w) : 0;                       // affected columns at the left end of the box
    const int pr = k - 1 - pl;
    const int nr = w0 + box_w > W - pr ? min(w0 + box_w - (W - pr), box_w) : 0;      // ... and at the right end
    const int ne = nl + nr;
    if (ne == 0) return;
    const int ncand = ne * box_h * 8;
    for (int idx = lane; idx < ncand; idx += 32) {
        const int g = idx & 7, e = idx >> 3;
        const int hb = e / ne, j = e - hb * ne;
        const int wb = j < nl ? j : box_w - nr + (j - nl);
        const int p = w0 + wb;
        const int lo = max(0, pl - p) * C;                                  // window elements [0, lo) lie left of column 0
        const int hi = min(k, W - p + pl) * C;                              // ... [hi, k*C) right of column W-1
        const int el = c0 + 8 * g;
        if (el < lo || el >= hi) {
            const int r = hb * box_w + wb;
            const uint32_t addr = tile + (uint32_t)r * 128u + (uint32_t)((g ^ (r & 7)) << 4);
            asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(addr), "r"(0u) : "memory");
        }
    }
    fence_proxy_async();
}

static constexpr int TC_THREADS = 192;        // 6 warps: TMA, MMA, 4 x epilogue
// conv_tc_kernel: 10 warps = TMA, MMA, 8 x epilogue.  The epilogue of a 32-column chunk (tcgen05.ld, bf16 pack, stores,
// instance-norm column sums) costs a lone warp ~2200 cycles of mostly exposed latency, more than the main loop of every
// layer but the 9-tap trunk convs; two warps per TMEM lane quarter (each takes every other chunk) hide each other's latency.
static constexpr int CONV_THREADS = 320;
static constexpr int EPI_WARPS = 8;
static constexpr int EPI_SCRATCH = EPI_WARPS * 32 * 17 * 4;      // per warp: a [32 rows][16 columns + 1] float transpose buffer
static constexpr int A_TILE_BYTES = 128 * 128;  // 128 rows x 64 bf16
static constexpr int TMEM_COLS = 512;

// ------------------------------------------------------------------------------------------
// forward / data-gradient kernel
// ------------------------------------------------------------------------------------------
// BK16 (the 16-channel-group SWIZZLE_32B mode of the U-Net layers) is a template parameter: as a run-time branch it cost
// the 64-channel instance 19 % (1216 -> 987 TFLOP/s at the C3 trunk shape, same box, back to back).
// DUAL: the variant for N tiles <= 128 columns, two CTAs per SM (6 warps, <= 110 KB of shared memory and 256 TMEM columns
// each).  With few K steps per tile the single MMA-issuing thread's chain (mbarrier wait -> MMAs -> commit, ~600 cycles
// per K step; ~2100 cycles per 3-step tile even with loads and epilogue switched off) bounds the SM, not the tensor pipe;
// two resident CTAs overlap their chains.
// WIN: window mode (TcConvArgs::win_*): the MMA warp patches the image-row edges of every A tile before issuing.
// EPI: epilogue variant.  0 = as described above.
//   1 = TWO groups of epilogue warps, one per TMEM accumulator (18 warps, or 10 per CTA with DUAL): tiles alternate between
//       the groups, so the epilogue of tile i+1 starts while tile i's is still running.  For the layers whose main loop is
//       shorter than the epilogue of a tile (few K steps, narrow N), where the epilogue warps are the critical path.
// (A low-shared-memory variant -- rows stored straight from registers, statistics by warp shuffles -- was measured slower
// on the trunk convs: forward 4.43 -> 4.70 ms, fold-mode data gradient 4.81 -> 7.92 ms per step.)
template <bool BK16, bool DUAL, bool WIN = false, int EPI = 0>
__global__ void __launch_bounds__((DUAL ? TC_THREADS : CONV_THREADS) + (EPI == 1 ? (DUAL ? 128 : 256) : 0), DUAL ? 2 : 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
               bf16* __restrict__ out, const float* __restrict__ bias, const TcConvArgs a) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // bk16: a K step holds `kg` 16-channel groups = `groups` channel groups of one tap, or (tps > 1) all cin16 groups of tps taps
    const uint32_t kg = BK16 ? (uint32_t)(a.tps > 1 ? a.tps * a.cin16 : a.groups) : 0u;
    const uint32_t a_bytes = BK16 ? kg * 4096u : (uint32_t)A_TILE_BYTES;
    const uint32_t stage_bytes = BK16 ? kg * (4096u + (uint32_t)a.bn * 32u) : A_TILE_BYTES + (uint32_t)a.bn * 128u;
    const int S = a.stages;
    const uint32_t bar0 = smem0 + S * stage_bytes;          // full[S], empty[S], tfull[2], tempty[2], tmem ptr
    auto full = [&](int s) { return bar0 + 8u * s; };
    auto empty = [&](int s) { return bar0 + 8u * (S + s); };
    auto tfull = [&](int i) { return bar0 + 8u * (2 * S + i); };
    auto tempty = [&](int i) { return bar0 + 8u * (2 * S + 2 + i); };
    const uint32_t tmem_slot = bar0 + 8u * (2 * S + 4);
    volatile uint32_t* tmem_slot_ptr = (volatile uint32_t*)(smem_raw + (tmem_slot - smem_u32(smem_raw)));
    float* stat_sm = reinterpret_cast<float*>(smem_raw + (bar0 + 256u - smem_u32(smem_raw)));   // epilogue: per-warp transpose buffers

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&mapA);
        tma_prefetch_desc(&mapB);
        for (int s = 0; s < S; ++s) { mbar_init(full(s), 1); mbar_init(empty(s), 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(tfull(i), 1); mbar_init(tempty(i), DUAL ? 4 : EPI_WARPS); }
        fence_barrier_init();
    }
    constexpr uint32_t ACC_COLS = DUAL ? 128u : 256u;      // TMEM columns per accumulator (two accumulators)
    if (warp == 1) tmem_alloc(tmem_slot, 2 * ACC_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    griddep_launch();        // PDL: the next kernel may take the slots this grid frees and run its prologue
    griddep_wait();          // everything above is independent of the previous kernel; global memory is not

    const int total_tiles = a.nb * a.tiles_per_img * a.n_blocks_n;
    // bk16: cchunks = K steps per tap (blocks of `groups` 16-channel groups), or several taps per K step
    const int ksteps = (BK16 && a.tps > 1) ? (a.n_taps + a.tps - 1) / a.tps : a.n_taps * a.cchunks;

    if (warp == 0) {        // TMA producer: the whole warp runs the loop (uniform control flow), one elected lane issues
        int s = 0; uint32_t ph = 0;
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
            const int nblk = t % a.n_blocks_n, mt = t / a.n_blocks_n;
            const int img = a.n0 + mt / a.tiles_per_img, ti = mt % a.tiles_per_img;
            const int w0 = (ti % a.tiles_w) * a.Wb, h0 = (ti / a.tiles_w) * a.Hb;
            for (int ks = 0; ks < ksteps; ++ks) {
                if constexpr (BK16) {
                    if (a.tps > 1) {        // Cin <= 64: `tps` taps per K step, one A box each, ONE B box of tps*bn weight rows
                        const int t0 = ks * a.tps, nt = min(a.tps, a.n_taps - t0);
                        mbar_wait(empty(s), ph ^ 1u);
                        const uint32_t sa = smem0 + s * stage_bytes;
                        if (elect_one()) {
                            mbar_expect_tx(full(s), (uint32_t)(nt * a.cin16) * 4096u + kg * (uint32_t)a.bn * 32u);
                            for (int j = 0; j < nt; ++j)
                                tma_load_5d(sa + (uint32_t)(j * a.cin16) * 4096u, &mapA, full(s), 0, w0 + a.dw[t0 + j], h0 + a.dh[t0 + j],
                                            0, img);
                            tma_load_3d(sa + a_bytes, &mapB, full(s), 0, a.tb[t0] * a.b_rows_per_tap, 0);
                        }
                        __syncwarp();
                        if (++s == S) { s = 0; ph ^= 1u; }
                        continue;
                    }
                }
                const int tap = ks / a.cchunks, cc = ks - tap * a.cchunks;
                mbar_wait(empty(s), ph ^ 1u);
                const uint32_t sa = smem0 + s * stage_bytes;
                if (elect_one()) {
                    mbar_expect_tx(full(s), stage_bytes);
                    if constexpr (BK16) {       // one box = `groups` 16-channel groups of this tap (out-of-range groups are zero-filled)
                        tma_load_5d(sa, &mapA, full(s), 0, w0 + a.dw[tap], h0 + a.dh[tap], cc * a.groups, img);
                        tma_load_3d(sa + a_bytes, &mapB, full(s), 0, a.tb[tap] * a.b_rows_per_tap + nblk * a.bn, cc * a.groups);
                    } else {
                        tma_load_5d(sa, &mapA, full(s), cc * 64 + a.dc[tap], w0 + a.dw[tap], a.dp[tap], h0 + a.dh[tap], img);
                        tma_load_2d(sa + A_TILE_BYTES, &mapB, full(s), cc * 64, a.tb[tap] * a.b_rows_per_tap + nblk * a.bn);
                    }
                }
                __syncwarp();
                if (++s == S) { s = 0; ph ^= 1u; }
            }
        }
    } else if (warp == 1) {
        if constexpr (WIN) {        // all 32 lanes: wait, patch the row edges of the A tile, then lane 0 issues
            int s = 0; uint32_t ph = 0;
            int it = 0;
            for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
                const int acc = it & 1;
                const uint32_t acc_ph = (uint32_t)(it >> 1) & 1u;
                const int ti = (t / a.n_blocks_n) % a.tiles_per_img;
                const int w0 = (ti % a.tiles_w) * a.Wb;
                mbar_wait(tempty(acc), acc_ph ^ 1u);
                const uint32_t d_tmem = tmem_base + (uint32_t)acc * ACC_COLS;
                for (int ks = 0; ks < ksteps; ++ks) {
                    mbar_wait(full(s), ph);
                    const uint32_t sa = smem0 + s * stage_bytes;
                    win_fix(sa, a.Wb, a.Hb, w0, a.win_W, a.win_C, a.win_k, a.win_pl, a.dc[ks], lane);
                    __syncwarp();
                    tc_fence_after();
                    if (elect_one()) {
                        const uint64_t adesc = make_smem_desc(sa, 16, 1024);
                        const uint64_t bdesc = make_smem_desc(sa + A_TILE_BYTES, 16, 1024);
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            umma_bf16(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), a.idesc,
                                      (uint32_t)((ks | k) != 0));
                        umma_commit(empty(s));
                        if (ks == ksteps - 1) umma_commit(tfull(acc));
                    }
                    __syncwarp();
                    if (++s == S) { s = 0; ph ^= 1u; }
                }
            }
        } else {        // MMA issuer: uniform control flow for the whole warp, one elected lane issues (see elect_one)
            int s = 0; uint32_t ph = 0;
            int it = 0;
            for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
                const int acc = it & 1;
                const uint32_t acc_ph = (uint32_t)(it >> 1) & 1u;
                mbar_wait(tempty(acc), acc_ph ^ 1u);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)acc * ACC_COLS;
                for (int ks = 0; ks < ksteps; ++ks) {
                    mbar_wait(full(s), ph);
                    tc_fence_after();
                    const uint32_t sa = smem0 + s * stage_bytes;
                    if (elect_one()) {
                        if constexpr (BK16) {       // one K = 16 MMA per 16-channel group
                            if (a.tps > 1) {
                                const int nt = min(a.tps, a.n_taps - ks * a.tps);
                                for (int j = 0; j < nt; ++j)
                                    for (int cg = 0; cg < a.cin16; ++cg)
                                        umma_bf16(d_tmem, make_smem_desc32(sa + (uint32_t)(j * a.cin16 + cg) * 4096u),
                                                  make_smem_desc32(sa + a_bytes + (uint32_t)(cg * a.tps + j) * (uint32_t)a.bn * 32u), a.idesc,
                                                  (uint32_t)((ks | j | cg) != 0));
                            } else {
                                const int cc = ks % a.cchunks;
                                const int ng = min(a.groups, a.cin16 - cc * a.groups);
                                for (int gq = 0; gq < ng; ++gq)
                                    umma_bf16(d_tmem, make_smem_desc32(sa + (uint32_t)gq * 4096u),
                                              make_smem_desc32(sa + a_bytes + (uint32_t)gq * (uint32_t)a.bn * 32u), a.idesc,
                                              (uint32_t)((ks | gq) != 0));
                            }
                        } else {
                            const uint64_t adesc = make_smem_desc(sa, 16, 1024);
                            const uint64_t bdesc = make_smem_desc(sa + A_TILE_BYTES, 16, 1024);
#pragma unroll
                            for (int k = 0; k < 4; ++k)     // 4 x (K = 16 bf16 = 32 bytes) inside the 128-byte swizzle row
                                umma_bf16(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), a.idesc,
                                          (uint32_t)((ks | k) != 0));
                        }
                        umma_commit(empty(s));
                        if (ks == ksteps - 1) umma_commit(tfull(acc));
                    }
                    __syncwarp();
                    if (++s == S) { s = 0; ph ^= 1u; }
                }
            }
        }
    } else {
        const int q = warp & 3;                   // TMEM lane quarter this warp may read
        const int half = DUAL ? 0 : ((warp - 2) >> 2) & 1;   // two warps per quarter: even / odd 32-column chunks (DUAL: one warp, all chunks)
        const int grp = EPI == 1 ? (warp - 2) / (DUAL ? 4 : EPI_WARPS) : 0;      // EPI 1: this warp's accumulator
        float* tr = stat_sm + (warp - 2) * (32 * 17);
        int it = 0;
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
            const int acc = it & 1;
            if (EPI == 1 && acc != grp) continue;
            const uint32_t acc_ph = (uint32_t)(it >> 1) & 1u;
            const int nblk = t % a.n_blocks_n, mt = t / a.n_blocks_n;
            const int img = mt / a.tiles_per_img, ti = mt % a.tiles_per_img;      // img relative to the output base
            const int w0 = (ti % a.tiles_w) * a.Wb, h0 = (ti / a.tiles_w) * a.Hb;
            const int lin = h0 * a.out_P + w0 + q * 32 + lane;
            const int oh = lin / a.out_P, ow = lin - oh * a.out_P;
            const bool valid = ow < a.out_wvalid && oh < a.out_hvalid;
            bf16* dst = out + (((size_t)img * a.out_H + (oh * a.out_sy + a.out_oy)) * a.out_W + (ow * a.out_sx + a.out_ox)) * a.Cout +
                        (size_t)nblk * a.bn;
            if (a.fold_pad > 0) {                 // interior pixel of a reflection-padded grid -> the unpadded gradient
                const int fp = a.fold_pad, Hi = a.out_H - 2 * fp, Wi = a.out_W - 2 * fp;
                if (oh >= fp && oh < fp + Hi && ow >= fp && ow < fp + Wi)
                    dst = reinterpret_cast<bf16*>(reinterpret_cast<unsigned long long>(
                              a.out2 + (((size_t)img * Hi + (oh - fp)) * Wi + (ow - fp)) * a.Cout + (size_t)nblk * a.bn) |
                          (unsigned long long)(a.fold_acc != 0));      // bit 0 of the (16-byte aligned) pointer = accumulate
            }
            bf16* rowptr[4];                      // output rows this lane stores: row (lane/4 + 8i) of the warp's 32, or null
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const unsigned long long pv = __shfl_sync(0xffffffffu, valid ? (unsigned long long)dst : 0ull, (lane >> 2) + 8 * i);
                rowptr[i] = reinterpret_cast<bf16*>(pv);
            }
            mbar_wait(tfull(acc), acc_ph);
            tc_fence_after();
            const uint32_t taddr = tmem_base + (uint32_t)acc * ACC_COLS + ((uint32_t)(q * 32) << 16);
            for (int c0 = half * 32; c0 < a.bn; c0 += (DUAL ? 32 : 64)) {
                uint32_t v[32];
                tmem_ld32(taddr + (uint32_t)c0, v);
                if (a.dbg & 1) continue;
                const int cols = (BK16 || WIN) ? min(32, a.bn - c0) : 32;          // 16 when the N tile is not a multiple of 32
                if (bias) {
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (j < cols) v[j] = __float_as_uint(__uint_as_float(v[j]) + __ldg(bias + nblk * a.bn + c0 + j));
                }
                {
                    // a lane owns one pixel row (64 bytes of this chunk): storing it directly makes every warp store touch 32
                    // different lines with 16 bytes each.  Four 16-byte pieces per lane go through an XOR-swizzled 2 KB
                    // staging tile instead, and each store instruction writes the whole 64-byte segments of 8 rows.
                    uint4* stg = reinterpret_cast<uint4*>(tr);
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        stg[lane * 4 + (j ^ ((lane >> 1) & 3))] =
                            make_uint4(pack_bf16x2(__uint_as_float(v[8 * j]), __uint_as_float(v[8 * j + 1])),
                                       pack_bf16x2(__uint_as_float(v[8 * j + 2]), __uint_as_float(v[8 * j + 3])),
                                       pack_bf16x2(__uint_as_float(v[8 * j + 4]), __uint_as_float(v[8 * j + 5])),
                                       pack_bf16x2(__uint_as_float(v[8 * j + 6]), __uint_as_float(v[8 * j + 7])));
                    __syncwarp();
                    const int g = lane & 3;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int r = (lane >> 2) + 8 * i;
                        uint4 val = stg[r * 4 + (g ^ ((r >> 1) & 3))];
                        if (rowptr[i] && g * 8 < cols) {
                            const unsigned long long pv = reinterpret_cast<unsigned long long>(rowptr[i]);
                            uint4* gp = reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(pv & ~1ull) + c0 + g * 8);
                            if (pv & 1ull) {          // accumulate onto the gradient already there (the skip path wrote it)
                                const uint4 old = *gp;
                                val.x = add_bf16x2(val.x, old.x); val.y = add_bf16x2(val.y, old.y);
                                val.z = add_bf16x2(val.z, old.z); val.w = add_bf16x2(val.w, old.w);
                            }
                            *gp = val;
                        }
                    }
                    __syncwarp();
                }
                if (a.stats && !(a.dbg & 2)) {
                    // fused instance-norm statistics: per-channel sum and sum of squares over this warp's 32 rows, 16
                    // columns per pass through a padded shared-memory transpose (conflict-free both ways): lanes 0-15 sum
                    // rows 0-15 of column `lane`, lanes 16-31 rows 16-31 of column `lane-16`; one shuffle joins the halves
                    // and ONE 128-byte reduction adds {sum, sumsq} x 16 columns to the [img][channel][2] table
                    float* sp = a.stats + ((size_t)img * a.Cout + (size_t)nblk * a.bn + c0) * 2;
#pragma unroll
                    for (int pass = 0; pass < 2; ++pass) {
                        if (pass * 16 < cols) {
#pragma unroll
                            for (int j = 0; j < 16; ++j) tr[lane * 17 + j] = valid ? __uint_as_float(v[pass * 16 + j]) : 0.f;
                            __syncwarp();
                            const float* col = tr + (lane >> 4) * 16 * 17 + (lane & 15);
                            float s1 = 0.f, s2 = 0.f, s1b = 0.f, s2b = 0.f;
#pragma unroll
                            for (int r = 0; r < 16; r += 2) {
                                const float x0 = col[r * 17], x1 = col[(r + 1) * 17];
                                s1 += x0; s2 = fmaf(x0, x0, s2);
                                s1b += x1; s2b = fmaf(x1, x1, s2b);
                            }
                            s1 += s1b; s2 += s2b;
                            s1 += __shfl_xor_sync(0xffffffffu, s1, 16);
                            s2 += __shfl_xor_sync(0xffffffffu, s2, 16);
                            atomicAdd(sp + pass * 32 + 2 * (lane & 15) + (lane >> 4), (lane >> 4) ? s2 : s1);
                            __syncwarp();
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty(acc));
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 2 * ACC_COLS);
    }
}

// ------------------------------------------------------------------------------------------
// Window-form convolution with vertical halo reuse (the U-Net layers: stride-1 'same' convs, C % 8 == 0, N <= 256).
// conv_tc_kernel<WIN> loads one 16 KB A tile per (kernel row kh, window chunk j): k*k times the unique input bytes go
// through the L2 -> shared-memory path, and these narrow-N layers sit on its cap (measured ~11.4 TB/s of L2 reads against
// ~12 TB/s, 0.5 us per K step whatever N is).  Here the M tile is an 8 x 16 pixel box, a stage holds the (8 + k - 1) x 16
// HALO of one window chunk plus the k weight tiles of that chunk (rows (kh*nch + j)*N of the packed [kh][j][N][64] matrix), and the k kernel rows are k shifted views of the same
// shared-memory tile (descriptor start + kh * 16 rows): (8 + k - 1) / (8 k) of the A traffic (0.34 for k = 4, 0.25 for
// k = 7) and one barrier round trip per k K-steps.  Same warp roles and epilogue as conv_tc_kernel.
// ------------------------------------------------------------------------------------------
// (A second group of four epilogue warps for the DUAL instance -- one group per TMEM accumulator, as in conv_tc_kernel<EPI 1> --
// was measured neutral on C2 and slower on C5 (7.73 -> 8.23 ms per 32-image call): these layers are not epilogue-bound.)
template <bool DUAL>
__global__ void __launch_bounds__(DUAL ? TC_THREADS : CONV_THREADS, DUAL ? 2 : 1)
convw_tc_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                bf16* __restrict__ out, const float* __restrict__ bias, const TcConvArgs a) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int k = a.win_k, nch = a.cchunks, halo_h = a.Hb + k - 1;
    const uint32_t a_bytes = (uint32_t)(halo_h * a.Wb) * 128u, b_tile = (uint32_t)a.bn * 128u;
    const uint32_t stage_bytes = a_bytes + (uint32_t)k * b_tile;
    const int S = a.stages;
    const uint32_t bar0 = smem0 + S * stage_bytes;
    auto full = [&](int s) { return bar0 + 8u * s; };
    auto empty = [&](int s) { return bar0 + 8u * (S + s); };
    auto tfull = [&](int i) { return bar0 + 8u * (2 * S + i); };
    auto tempty = [&](int i) { return bar0 + 8u * (2 * S + 2 + i); };
    const uint32_t tmem_slot = bar0 + 8u * (2 * S + 4);
    volatile uint32_t* tmem_slot_ptr = (volatile uint32_t*)(smem_raw + (tmem_slot - smem_u32(smem_raw)));
    float* stat_sm = reinterpret_cast<float*>(smem_raw + (bar0 + 256u - smem_u32(smem_raw)));

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&mapA);
        tma_prefetch_desc(&mapB);
        for (int s = 0; s < S; ++s) { mbar_init(full(s), 1); mbar_init(empty(s), 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(tfull(i), 1); mbar_init(tempty(i), DUAL ? 4 : EPI_WARPS); }
        fence_barrier_init();
    }
    constexpr uint32_t ACC_COLS = DUAL ? 128u : 256u;
    if (warp == 1) tmem_alloc(tmem_slot, 2 * ACC_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    griddep_launch();
    griddep_wait();

    const int total_tiles = a.nb * a.tiles_per_img;

    if (warp == 0) {        // uniform control flow for the whole warp, one elected lane issues (see elect_one)
        int s = 0; uint32_t ph = 0;
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
            const int img = a.n0 + t / a.tiles_per_img, ti = t % a.tiles_per_img;
            const int w0 = (ti % a.tiles_w) * a.Wb, h0 = (ti / a.tiles_w) * a.Hb;
            for (int j = 0; j < nch; ++j) {
                mbar_wait(empty(s), ph ^ 1u);
                const uint32_t sa = smem0 + s * stage_bytes;
                if (elect_one()) {
                    mbar_expect_tx(full(s), stage_bytes);
                    tma_load_5d(sa, &mapA, full(s), 64 * j, w0, 0, h0 - a.win_pt, img);
                    for (int kh = 0; kh < k; ++kh)
                        tma_load_2d(sa + a_bytes + (uint32_t)kh * b_tile, &mapB, full(s), 0, (kh * nch + j) * a.bn);
                }
                __syncwarp();
                if (++s == S) { s = 0; ph ^= 1u; }
            }
        }
    } else if (warp == 1) {
        int s = 0; uint32_t ph = 0;
        int it = 0;
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
            const int acc = it & 1;
            const uint32_t acc_ph = (uint32_t)(it >> 1) & 1u;
            const int ti = t % a.tiles_per_img;
            const int w0 = (ti % a.tiles_w) * a.Wb;
            mbar_wait(tempty(acc), acc_ph ^ 1u);
            const uint32_t d_tmem = tmem_base + (uint32_t)acc * ACC_COLS;
            for (int j = 0; j < nch; ++j) {
                mbar_wait(full(s), ph);
                const uint32_t sa = smem0 + s * stage_bytes;
                win_fix(sa, a.Wb, halo_h, w0, a.win_W, a.win_C, k, a.win_pl, 64 * j, lane);
                __syncwarp();
                tc_fence_after();
                if (elect_one()) {
                    for (int kh = 0; kh < k; ++kh) {
                        const uint64_t adesc = make_smem_desc(sa + (uint32_t)(kh * a.Wb) * 128u, 16, 1024);
                        const uint64_t bdesc = make_smem_desc(sa + a_bytes + (uint32_t)kh * b_tile, 16, 1024);
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk)
                            umma_bf16(d_tmem, adesc + (uint64_t)(kk * 2), bdesc + (uint64_t)(kk * 2), a.idesc,
                                      (uint32_t)((j | kh | kk) != 0));
                    }
                    umma_commit(empty(s));
                    if (j == nch - 1) umma_commit(tfull(acc));
                }
                __syncwarp();
                if (++s == S) { s = 0; ph ^= 1u; }
            }
        }
    } else {
        const int q = warp & 3;
        const int half = (warp - 2) >> 2;
        float* tr = stat_sm + (warp - 2) * (32 * 17);
        int it = 0;
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
            const int acc = it & 1;
            const uint32_t acc_ph = (uint32_t)(it >> 1) & 1u;
            const int img = t / a.tiles_per_img, ti = t % a.tiles_per_img;          // img relative to the output base
            const int w0 = (ti % a.tiles_w) * a.Wb, h0 = (ti / a.tiles_w) * a.Hb;
            const int rr = q * 32 + lane;                                         // tile row = (rr / Wb, rr % Wb) of the pixel box
            const int oh = h0 + rr / a.Wb, ow = w0 + rr % a.Wb;
            bf16* dst = out + (((size_t)img * a.out_H + oh) * a.out_W + ow) * a.Cout;
            bf16* rowptr[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const unsigned long long pv = __shfl_sync(0xffffffffu, (unsigned long long)dst, (lane >> 2) + 8 * i);
                rowptr[i] = reinterpret_cast<bf16*>(pv);
            }
            mbar_wait(tfull(acc), acc_ph);
            tc_fence_after();
            const uint32_t taddr = tmem_base + (uint32_t)acc * ACC_COLS + ((uint32_t)(q * 32) << 16);
            for (int c0 = half * 32; c0 < a.bn; c0 += (DUAL ? 32 : 64)) {
                uint32_t v[32];
                tmem_ld32(taddr + (uint32_t)c0, v);
                const int cols = min(32, a.bn - c0);
                if (bias) {
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (j < cols) v[j] = __float_as_uint(__uint_as_float(v[j]) + __ldg(bias + c0 + j));
                }
                {
                    uint4* stg = reinterpret_cast<uint4*>(tr);
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        stg[lane * 4 + (j ^ ((lane >> 1) & 3))] =
                            make_uint4(pack_bf16x2(__uint_as_float(v[8 * j]), __uint_as_float(v[8 * j + 1])),
                                       pack_bf16x2(__uint_as_float(v[8 * j + 2]), __uint_as_float(v[8 * j + 3])),
                                       pack_bf16x2(__uint_as_float(v[8 * j + 4]), __uint_as_float(v[8 * j + 5])),
                                       pack_bf16x2(__uint_as_float(v[8 * j + 6]), __uint_as_float(v[8 * j + 7])));
                    __syncwarp();
                    const int g = lane & 3;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int r = (lane >> 2) + 8 * i;
                        const uint4 val = stg[r * 4 + (g ^ ((r >> 1) & 3))];
                        if (g * 8 < cols) *reinterpret_cast<uint4*>(rowptr[i] + c0 + g * 8) = val;
                    }
                    __syncwarp();
                }
                if (a.stats) {
                    float* sp = a.stats + ((size_t)img * a.Cout + c0) * 2;
#pragma unroll
                    for (int pass = 0; pass < 2; ++pass) {
                        if (pass * 16 < cols) {
#pragma unroll
                            for (int j = 0; j < 16; ++j) tr[lane * 17 + j] = __uint_as_float(v[pass * 16 + j]);
                            __syncwarp();
                            const float* col = tr + (lane >> 4) * 16 * 17 + (lane & 15);
                            float s1 = 0.f, s2 = 0.f, s1b = 0.f, s2b = 0.f;
#pragma unroll
                            for (int r = 0; r < 16; r += 2) {
                                const float x0 = col[r * 17], x1 = col[(r + 1) * 17];
                                s1 += x0; s2 = fmaf(x0, x0, s2);
                                s1b += x1; s2b = fmaf(x1, x1, s2b);
                            }
                            s1 += s1b; s2 += s2b;
                            s1 += __shfl_xor_sync(0xffffffffu, s1, 16);
                            s2 += __shfl_xor_sync(0xffffffffu, s2, 16);
                            atomicAdd(sp + pass * 32 + 2 * (lane & 15) + (lane >> 4), (lane >> 4) ? s2 : s1);
                            __syncwarp();
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty(acc));
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 2 * ACC_COLS);
    }
}

// ------------------------------------------------------------------------------------------
// 2-CTA variant (cta_group::2): a CTA pair (one TPC) computes a 256-pixel x BN tile.  Each CTA loads ITS 128 pixel rows
// of A and HALF of the B rows; the leader's single thread issues tcgen05.mma.cta_group::2 (M = 256), which reads A/B
// halves from both CTAs' shared memory and writes 128 accumulator rows into each CTA's TMEM.  Per SM this halves the
// B bytes written by TMA and read by the MMA -- the 1-CTA kernel above is shared-memory-bandwidth bound
// (48 KB written + 48 KB read per 512 MMA cycles = 187 B/clk against ~128 B/clk).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_5d_2sm(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                                                int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar & 0xFEFFFFFFu), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar & 0xFEFFFFFFu), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t dst_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma_bf16_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                              uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive (once) on the barrier at the same offset in BOTH CTAs of the pair when the MMAs issued so far have completed
__device__ __forceinline__ void umma_commit_2sm(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint32_t bar) {      // arrive on CTA 0's copy of `bar`
    asm volatile(
        "{\n\t.reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, 0;\n\t"
        "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}"
        ::"r"(bar) : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC_THREADS, 1)
conv_tc2_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                bf16* __restrict__ out, const float* __restrict__ bias, const TcConvArgs a) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const uint32_t b_half = (uint32_t)(a.bn / 2) * 128u;
    const uint32_t stage_bytes = A_TILE_BYTES + b_half;
    const int S = a.stages;
    const uint32_t bar0 = smem0 + S * stage_bytes;
    auto full = [&](int s) { return bar0 + 8u * s; };
    auto empty = [&](int s) { return bar0 + 8u * (S + s); };
    auto tfull = [&](int i) { return bar0 + 8u * (2 * S + i); };
    auto tempty = [&](int i) { return bar0 + 8u * (2 * S + 2 + i); };
    const uint32_t tmem_slot = bar0 + 8u * (2 * S + 4);
    volatile uint32_t* tmem_slot_ptr = (volatile uint32_t*)(smem_raw + (tmem_slot - smem_u32(smem_raw)));

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&mapA);
        tma_prefetch_desc(&mapB);
        for (int s = 0; s < S; ++s) { mbar_init(full(s), 1); mbar_init(empty(s), 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(tfull(i), 1); mbar_init(tempty(i), 8); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc_2sm(tmem_slot, TMEM_COLS);
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    const int m_pairs = a.nb * a.tiles_per_img / 2;              // host guarantees an even number of M tiles
    const int total_pairs = m_pairs * a.n_blocks_n;
    const int n_clusters = gridDim.x / 2, cluster_id = blockIdx.x / 2;
    const int ksteps = a.n_taps * a.cchunks;

    if (warp == 0) {
        if (lane == 0) {
            int s = 0; uint32_t ph = 0;
            for (int p = cluster_id; p < total_pairs; p += n_clusters) {
                const int nblk = p % a.n_blocks_n, mt = 2 * (p / a.n_blocks_n) + (int)rank;
                const int img = a.n0 + mt / a.tiles_per_img, ti = mt % a.tiles_per_img;
                const int w0 = (ti % a.tiles_w) * a.Wb, h0 = (ti / a.tiles_w) * a.Hb;
                for (int ks = 0; ks < ksteps; ++ks) {
                    const int tap = ks / a.cchunks, cc = ks - tap * a.cchunks;
                    mbar_wait(empty(s), ph ^ 1u);
                    if (leader) mbar_expect_tx(full(s), 2u * stage_bytes);
                    const uint32_t sa = smem0 + s * stage_bytes;
                    tma_load_5d_2sm(sa, &mapA, full(s), cc * 64 + a.dc[tap], w0 + a.dw[tap], a.dp[tap], h0 + a.dh[tap], img);
                    tma_load_2d_2sm(sa + A_TILE_BYTES, &mapB, full(s), cc * 64,
                                    a.tb[tap] * a.b_rows_per_tap + nblk * a.bn + (int)rank * (a.bn / 2));
                    if (++s == S) { s = 0; ph ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && leader) {
            int s = 0; uint32_t ph = 0;
            int it = 0;
            for (int p = cluster_id; p < total_pairs; p += n_clusters, ++it) {
                const int acc = it & 1;
                const uint32_t acc_ph = (uint32_t)(it >> 1) & 1u;
                mbar_wait(tempty(acc), acc_ph ^ 1u);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)acc * 256u;
                for (int ks = 0; ks < ksteps; ++ks) {
                    mbar_wait(full(s), ph);
                    tc_fence_after();
                    const uint32_t sa = smem0 + s * stage_bytes;
                    const uint64_t adesc = make_smem_desc(sa, 16, 1024);
                    const uint64_t bdesc = make_smem_desc(sa + A_TILE_BYTES, 16, 1024);
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_bf16_2sm(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), a.idesc,
                                      (uint32_t)((ks | k) != 0));
                    umma_commit_2sm(empty(s));
                    if (++s == S) { s = 0; ph ^= 1u; }
                }
                umma_commit_2sm(tfull(acc));
            }
        }
    } else {
        const int q = warp & 3;
        int it = 0;
        for (int p = cluster_id; p < total_pairs; p += n_clusters, ++it) {
            const int acc = it & 1;
            const uint32_t acc_ph = (uint32_t)(it >> 1) & 1u;
            const int nblk = p % a.n_blocks_n, mt = 2 * (p / a.n_blocks_n) + (int)rank;
            const int img = mt / a.tiles_per_img, ti = mt % a.tiles_per_img;
            const int w0 = (ti % a.tiles_w) * a.Wb, h0 = (ti / a.tiles_w) * a.Hb;
            const int lin = h0 * a.out_P + w0 + q * 32 + lane;
            const int oh = lin / a.out_P, ow = lin - oh * a.out_P;
            const bool valid = ow < a.out_wvalid && oh < a.out_hvalid;
            bf16* dst = out + (((size_t)img * a.out_H + (oh * a.out_sy + a.out_oy)) * a.out_W + (ow * a.out_sx + a.out_ox)) * a.Cout +
                        (size_t)nblk * a.bn;
            mbar_wait(tfull(acc), acc_ph);
            tc_fence_after();
            const uint32_t taddr = tmem_base + (uint32_t)acc * 256u + ((uint32_t)(q * 32) << 16);
            for (int c0 = 0; c0 < a.bn; c0 += 32) {
                uint32_t v[32];
                tmem_ld32(taddr + (uint32_t)c0, v);
                if (a.dbg & 1) continue;
                if (valid) {
                    uint32_t pk[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        float x0 = __uint_as_float(v[2 * j]), x1 = __uint_as_float(v[2 * j + 1]);
                        if (bias) { x0 += __ldg(bias + nblk * a.bn + c0 + 2 * j); x1 += __ldg(bias + nblk * a.bn + c0 + 2 * j + 1); }
                        pk[j] = pack_bf16x2(x0, x1);
                    }
                    uint4* d4 = reinterpret_cast<uint4*>(dst + c0);
#pragma unroll
                    for (int j = 0; j < 4; ++j) d4[j] = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_leader(tempty(acc));
        }
    }
    tc_fence_before();
    cluster_sync_all();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc_2sm(tmem_base, TMEM_COLS);
    }
}

// ------------------------------------------------------------------------------------------
// weight-gradient kernel.  One CTA = one (tap, 128-row block of operand A, BN-column block of operand B, K split).
//   normal     : A = X  tiles (rows = input channels),  B = dY tiles (cols = output channels)
//   transposed : A = dY tiles (rows = output channels), B = X  tiles (cols = input channels)   [Cin = 64 layers]
// smem stage: [2 boxes of A (64ch x 64px)] [BN/64 boxes of B (64ch x 64px)], each box 8 KB, MN-major.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TC_THREADS, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapDY,
                float* __restrict__ dw, const TcWgradArgs a) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nb_boxes = a.bn / 64;
    const uint32_t stage_bytes = (uint32_t)(2 + nb_boxes) * 8192u;
    const int S = a.stages;
    const uint32_t bar0 = smem0 + S * stage_bytes;
    auto full = [&](int s) { return bar0 + 8u * s; };
    auto empty = [&](int s) { return bar0 + 8u * (S + s); };
    const uint32_t tfull = bar0 + 8u * (2 * S);
    const uint32_t tmem_slot = bar0 + 8u * (2 * S + 1);
    volatile uint32_t* tmem_slot_ptr = (volatile uint32_t*)(smem_raw + (tmem_slot - smem_u32(smem_raw)));

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&mapX);
        tma_prefetch_desc(&mapDY);
        for (int s = 0; s < S; ++s) { mbar_init(full(s), 1); mbar_init(empty(s), 1); }
        mbar_init(tfull, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 256);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    griddep_launch();        // PDL: the next kernel may take the slots this grid frees and run its prologue
    griddep_wait();          // everything above is independent of the previous kernel; global memory is not

    // decode the work unit
    int u = blockIdx.x;
    const int split = u % a.splits; u /= a.splits;
    const int bblk = u % a.b_blocks; u /= a.b_blocks;
    const int ablk = u % a.a_blocks; u /= a.a_blocks;
    const int tap = u;
    const int total_chunks = a.nb * a.chunks_per_img;
    const int per = (total_chunks + a.splits - 1) / a.splits;
    const int q_begin = split * per;
    const int q_end = min(total_chunks, q_begin + per);
    const int nq = q_end - q_begin;
    // channel bases of the X boxes and of the dY boxes in this unit
    const int x_c0 = a.transposed ? bblk * a.bn : ablk * 128;
    const int y_c0 = a.transposed ? ablk * 128 : bblk * a.bn;
    const int x_boxes = a.transposed ? nb_boxes : 2, y_boxes = a.transposed ? 2 : nb_boxes;
    const uint32_t x_off = a.transposed ? 2 * 8192u : 0u, y_off = a.transposed ? 0u : 2 * 8192u;

    if (warp == 0) {        // uniform control flow for the whole warp, one elected lane issues (see elect_one)
        int s = 0; uint32_t ph = 0;
        for (int q = q_begin; q < q_end; ++q) {
            const int img = q / a.chunks_per_img, r = q % a.chunks_per_img;
            const int w0 = (r % a.chunks_w) * a.Wk, h0 = (r / a.chunks_w) * a.Hk;
            mbar_wait(empty(s), ph ^ 1u);
            const uint32_t sa = smem0 + s * stage_bytes;
            if (elect_one()) {
                mbar_expect_tx(full(s), stage_bytes);
                // dense operands are fetched with ONE grouped box instead of one box per 64 channels
                if (a.x_grouped)
                    tma_load_5d(sa + x_off, &mapX, full(s), 0, w0 + a.dw[tap], h0 + a.dh[tap], x_c0 / 64, a.n0 + img);
                else if (a.stack2)      // 64-channel X: the two 64-row halves of the A tile are two different taps
                    for (int b = 0; b < 2; ++b)
                        tma_load_5d(sa + x_off + b * 8192u, &mapX, full(s), 0, w0 + a.dw[2 * tap + b], 0, h0 + a.dh[2 * tap + b],
                                    a.n0 + img);
                else
                    for (int b = 0; b < x_boxes; ++b)
                        tma_load_5d(sa + x_off + b * 8192u, &mapX, full(s), x_c0 + b * 64 + a.dc[tap], w0 + a.dw[tap], a.dp[tap],
                                    h0 + a.dh[tap], a.n0 + img);
                if (a.y_grouped)
                    tma_load_5d(sa + y_off, &mapDY, full(s), 0, w0 + a.dy_off, h0 + a.dy_off, y_c0 / 64, a.y_n0 + img);
                else
                    for (int b = 0; b < y_boxes; ++b)
                        tma_load_5d(sa + y_off + b * 8192u, &mapDY, full(s), y_c0 + b * 64, w0 + a.dy_off, 0, h0 + a.dy_off,
                                    a.y_n0 + img);
            }
            __syncwarp();
            if (++s == S) { s = 0; ph ^= 1u; }
        }
    } else if (warp == 1) {
        if (nq > 0) {
            int s = 0; uint32_t ph = 0;
            for (int i = 0; i < nq; ++i) {
                mbar_wait(full(s), ph);
                tc_fence_after();
                const uint32_t sa = smem0 + s * stage_bytes;
                if (elect_one()) {
                    // MN-major canonical layout: 64 channels contiguous (128 B), pixels (K) at 128 B, 8-pixel groups at
                    // SBO = 1024 B, next 64-channel group at LBO = 8192 B.  One MMA consumes K = 16 pixels = 2048 B.
                    const uint64_t adesc = make_smem_desc(sa, 8192, 1024);
                    const uint64_t bdesc = make_smem_desc(sa + 2 * 8192u, 8192, 1024);
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_bf16(tmem_base, adesc + (uint64_t)(k * 128), bdesc + (uint64_t)(k * 128), a.idesc,
                                  (uint32_t)((i | k) != 0));
                    umma_commit(empty(s));
                    if (i == nq - 1) umma_commit(tfull);
                }
                __syncwarp();
                if (++s == S) { s = 0; ph ^= 1u; }
            }
        }
    } else if (nq > 0) {
        const int q = warp & 3;
        mbar_wait(tfull, 0);
        tc_fence_after();
        const int row = ablk * 128 + q * 32 + lane;
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
        if (!a.transposed) {        // row = ci, columns = co (contiguous in dW): 16-byte vector reductions
            float* dst = dw + ((size_t)tap * a.Cin + row) * a.Cout + (size_t)bblk * a.bn;
            for (int c0 = 0; c0 < a.bn; c0 += 32) {
                uint32_t v[32];
                tmem_ld32(taddr + (uint32_t)c0, v);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + c0 + 4 * j),
                                 "f"(__uint_as_float(v[4 * j])), "f"(__uint_as_float(v[4 * j + 1])),
                                 "f"(__uint_as_float(v[4 * j + 2])), "f"(__uint_as_float(v[4 * j + 3]))
                                 : "memory");
                }
            }
        } else {                    // row = co (lanes -> consecutive addresses), columns = ci
            float* dst = dw + ((size_t)tap * a.Cin + (size_t)bblk * a.bn) * a.Cout + row;
            for (int c0 = 0; c0 < a.bn; c0 += 32) {
                uint32_t v[32];
                tmem_ld32(taddr + (uint32_t)c0, v);
#pragma unroll
                for (int j = 0; j < 32; ++j) atomicAdd(dst + (size_t)(c0 + j) * a.Cout, __uint_as_float(v[j]));
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 256);
    }
}

// ------------------------------------------------------------------------------------------
// weight gradient for 16-multiple channel counts (U-Net layers).  The GEMM M dimension is the flattened (tap, ci) index
// of TF's HWIO kernel, cut into blocks of 128 rows = 8 slots of 16 channels; every slot is its own TMA box of the
// activation (16 channels x 64 pixels, shifted by the slot's tap; zero padding = out-of-bounds fill), so one block may
// stack several taps.  B = the dY chunk (all Cout channels, one grouped box).  MN-major SWIZZLE_32B operands,
// K = 64 pixels per stage, split-K over pixels, fp32 reductions straight into dW[(tap*Cin+ci)][co].
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t make_smem_desc32_mn(uint32_t saddr, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;     // next 16-channel group
    d |= (uint64_t)(256 >> 4) << 32;                      // SBO: next 8-pixel group
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)6 << 61;                               // SWIZZLE_32B
    return d;
}

__global__ void __launch_bounds__(TC_THREADS, 1)
wgrad16_tc_kernel(const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapDY,
                  float* __restrict__ dw, const TcWgrad16Args a) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ygroups = a.Cout / 16;
    const uint32_t stage_bytes = (uint32_t)(8 + ygroups) * 2048u;
    const int S = a.stages;
    const uint32_t bar0 = smem0 + S * stage_bytes;
    auto full = [&](int s) { return bar0 + 8u * s; };
    auto empty = [&](int s) { return bar0 + 8u * (S + s); };
    const uint32_t tfull = bar0 + 8u * (2 * S);
    const uint32_t tmem_slot = bar0 + 8u * (2 * S + 1);
    volatile uint32_t* tmem_slot_ptr = (volatile uint32_t*)(smem_raw + (tmem_slot - smem_u32(smem_raw)));

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&mapX);
        tma_prefetch_desc(&mapDY);
        for (int s = 0; s < S; ++s) { mbar_init(full(s), 1); mbar_init(empty(s), 1); }
        mbar_init(tfull, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 256);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    griddep_launch();        // PDL: the next kernel may take the slots this grid frees and run its prologue
    griddep_wait();          // everything above is independent of the previous kernel; global memory is not

    const int split = blockIdx.x % a.splits, mb = blockIdx.x / a.splits;
    const int total_chunks = a.nb * a.chunks_per_img;
    const int per = (total_chunks + a.splits - 1) / a.splits;
    const int q_begin = split * per;
    const int q_end = min(total_chunks, q_begin + per);
    const int nq = q_end - q_begin;
    const int rows_total = a.n_taps * a.Cin;
    const int nslots = min(8, (rows_total - mb * 128 + 15) / 16);      // live 16-row slots of this block

    if (warp == 0) {        // uniform control flow for the whole warp, one elected lane issues (see elect_one)
        int s = 0; uint32_t ph = 0;
        for (int q = q_begin; q < q_end; ++q) {
            const int img = q / a.chunks_per_img, r = q % a.chunks_per_img;
            const int w0 = (r % a.chunks_w) * a.Wk, h0 = (r / a.chunks_w) * a.Hk;
            mbar_wait(empty(s), ph ^ 1u);
            const uint32_t sa = smem0 + s * stage_bytes;
            if (elect_one()) {
                mbar_expect_tx(full(s), (uint32_t)(nslots + ygroups) * 2048u);
                for (int j = 0; j < nslots; ++j) {
                    const int row0 = (mb * 8 + j) * 16;
                    const int tap = row0 / a.Cin, cg = (row0 - tap * a.Cin) / 16;
                    tma_load_5d(sa + (uint32_t)j * 2048u, &mapX, full(s), 0, w0 + a.dw[tap], h0 + a.dh[tap], cg, a.n0 + img);
                }
                tma_load_5d(sa + 8u * 2048u, &mapDY, full(s), 0, w0, h0, 0, a.y_n0 + img);
            }
            __syncwarp();
            if (++s == S) { s = 0; ph ^= 1u; }
        }
    } else if (warp == 1) {
        if (nq > 0) {
            int s = 0; uint32_t ph = 0;
            for (int i = 0; i < nq; ++i) {
                mbar_wait(full(s), ph);
                tc_fence_after();
                const uint32_t sa = smem0 + s * stage_bytes;
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < 4; ++k)       // K = 16 pixels = 512 bytes inside each 2 KB group
                        umma_bf16(tmem_base, make_smem_desc32_mn(sa + (uint32_t)k * 512u, 2048),
                                  make_smem_desc32_mn(sa + 8u * 2048u + (uint32_t)k * 512u, 2048), a.idesc, (uint32_t)((i | k) != 0));
                    umma_commit(empty(s));
                    if (i == nq - 1) umma_commit(tfull);
                }
                __syncwarp();
                if (++s == S) { s = 0; ph ^= 1u; }
            }
        }
    } else if (nq > 0) {
        const int q = warp & 3;
        mbar_wait(tfull, 0);
        tc_fence_after();
        const int row = mb * 128 + q * 32 + lane;
        const bool live = row < rows_total && (q * 32 + lane) < nslots * 16;
        float* dst = dw + (size_t)row * a.Cout;
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
        for (int c0 = 0; c0 < a.Cout; c0 += 32) {
            uint32_t v[32];
            tmem_ld32(taddr + (uint32_t)c0, v);
            const int cols = min(32, a.Cout - c0);
            if (live) {
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    if (j * 4 < cols)
                        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + c0 + 4 * j),
                                     "f"(__uint_as_float(v[4 * j])), "f"(__uint_as_float(v[4 * j + 1])),
                                     "f"(__uint_as_float(v[4 * j + 2])), "f"(__uint_as_float(v[4 * j + 3]))
                                     : "memory");
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 256);
    }
}

// ------------------------------------------------------------------------------------------
// weight gradient in window mode (U-Net layers, any C % 8 == 0): dW[kh][(kw, ci)][co] = sum_pixels Xwin[p][(kw, ci)] * dY[p][co].
// The A operand is the SAME window tile the forward conv loads (64 window elements x 64 pixels, MN-major here), two K
// steps of the forward conv stacked in the 128 MMA rows; B = the dY chunk as 16-channel groups (SWIZZLE_32B, one box).
// K = 64 pixels per stage, split over the pixel range, fp32 vector reductions straight into TF's HWIO gradient (the row
// index kh*k*C + e of the window element IS the HWIO row).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TC_THREADS, 1)
wgradw_tc_kernel(const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapDY,
                 float* __restrict__ dw, const TcWgradWArgs a) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ygroups = a.Cout / 16;
    // a unit = `halves` (2 or 4) window chunks = 1 or 2 accumulators of 128 rows that share ONE dY load and one barrier
    // round trip per 64-pixel stage (the per-stage chain of the issuing thread, not the tensor pipe, bounds these layers)
    const int halves = a.halves;
    const uint32_t a_bytes = (uint32_t)halves * 8192u;
    const uint32_t stage_bytes = a_bytes + (uint32_t)ygroups * 2048u;
    const int S = a.stages;
    const uint32_t bar0 = smem0 + S * stage_bytes;
    auto full = [&](int s) { return bar0 + 8u * s; };
    auto empty = [&](int s) { return bar0 + 8u * (S + s); };
    const uint32_t tfull = bar0 + 8u * (2 * S);
    const uint32_t tmem_slot = bar0 + 8u * (2 * S + 1);
    volatile uint32_t* tmem_slot_ptr = (volatile uint32_t*)(smem_raw + (tmem_slot - smem_u32(smem_raw)));
    const uint32_t acc_cols = a.Cout <= 128 ? 128u : 256u;      // column pitch of the accumulators (halves == 4 needs Cout <= 128)

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&mapX);
        tma_prefetch_desc(&mapDY);
        for (int s = 0; s < S; ++s) { mbar_init(full(s), 1); mbar_init(empty(s), 1); }
        mbar_init(tfull, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 256);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    griddep_launch();
    griddep_wait();

    const int split = blockIdx.x % a.splits, u = blockIdx.x / a.splits;
    const int total_chunks = a.nb * a.chunks_per_img;
    const int per = (total_chunks + a.splits - 1) / a.splits;
    const int q_begin = split * per;
    const int q_end = min(total_chunks, q_begin + per);
    const int nq = q_end - q_begin;
    const int step0 = halves * u;
    const int nhalf = min(halves, a.steps - step0);          // live 64-row halves of this unit
    const int npair = (nhalf + 1) / 2;

    if (warp == 0) {        // uniform control flow for the whole warp, one elected lane issues (see elect_one)
        int s = 0; uint32_t ph = 0;
        for (int q = q_begin; q < q_end; ++q) {
            const int img = q / a.chunks_per_img, r = q % a.chunks_per_img;
            const int w0 = (r % a.chunks_w) * a.Wk, h0 = (r / a.chunks_w) * a.Hk;
            mbar_wait(empty(s), ph ^ 1u);
            const uint32_t sa = smem0 + s * stage_bytes;
            if (elect_one()) {
                mbar_expect_tx(full(s), (uint32_t)nhalf * 8192u + (uint32_t)ygroups * 2048u);
                for (int b = 0; b < nhalf; ++b)
                    tma_load_5d(sa + (uint32_t)b * 8192u, &mapX, full(s), a.dc[step0 + b], w0, 0, h0 + a.dh[step0 + b], a.n0 + img);
                tma_load_5d(sa + a_bytes, &mapDY, full(s), 0, w0, h0, 0, a.y_n0 + img);
            }
            __syncwarp();
            if (++s == S) { s = 0; ph ^= 1u; }
        }
    } else if (warp == 1) {
        if (nq > 0) {
            int s = 0; uint32_t ph = 0;
            for (int i = 0; i < nq; ++i) {
                const int q = q_begin + i;
                const int w0 = ((q % a.chunks_per_img) % a.chunks_w) * a.Wk;
                mbar_wait(full(s), ph);
                const uint32_t sa = smem0 + s * stage_bytes;
                for (int b = 0; b < nhalf; ++b)
                    win_fix(sa + (uint32_t)b * 8192u, a.Wk, a.Hk, w0, a.W, a.C, a.k, a.pl, a.dc[step0 + b], lane);
                __syncwarp();
                tc_fence_after();
                if (elect_one()) {
                    // A: MN-major SWIZZLE_128B (64 elements contiguous, pixels at 128 B, 8-pixel groups at SBO = 1 KB, the second
                    // 64-row half at LBO = 8 KB); B: MN-major SWIZZLE_32B 16-channel groups at LBO = 2 KB.  K = 16 pixels per MMA.
                    for (int pi = 0; pi < npair; ++pi) {
                        const uint64_t adesc = make_smem_desc(sa + (uint32_t)pi * 16384u, 8192, 1024);
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            umma_bf16(tmem_base + (uint32_t)pi * acc_cols, adesc + (uint64_t)(k * 128),
                                      make_smem_desc32_mn(sa + a_bytes + (uint32_t)k * 512u, 2048), a.idesc, (uint32_t)((i | k) != 0));
                    }
                    umma_commit(empty(s));
                    if (i == nq - 1) umma_commit(tfull);
                }
                __syncwarp();
                if (++s == S) { s = 0; ph ^= 1u; }
            }
        }
    } else if (nq > 0) {
        const int q = warp & 3;
        mbar_wait(tfull, 0);
        tc_fence_after();
        const int row = q * 32 + lane, half = row >> 6, i = row & 63;
        for (int pi = 0; pi < npair; ++pi) {
            const int step = step0 + 2 * pi + half;
            bool live = step < a.steps;
            float* dst = dw;
            if (live) {
                const int kh = step / a.nch, j = step - kh * a.nch;
                const int e = 64 * j + i;
                const int kw = e / a.C, ci = e - kw * a.C;
                live = e < a.k * a.C && ci < a.Creal;
                dst = dw + ((size_t)(kh * a.k + kw) * a.Creal + ci) * a.Cout;
            }
            const uint32_t taddr = tmem_base + (uint32_t)pi * acc_cols + ((uint32_t)(q * 32) << 16);
            for (int c0 = 0; c0 < a.Cout; c0 += 32) {
                uint32_t v[32];
                tmem_ld32(taddr + (uint32_t)c0, v);
                const int cols = min(32, a.Cout - c0);
                if (live) {
#pragma unroll
                    for (int j4 = 0; j4 < 8; ++j4)
                        if (j4 * 4 < cols)
                            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + c0 + 4 * j4),
                                         "f"(__uint_as_float(v[4 * j4])), "f"(__uint_as_float(v[4 * j4 + 1])),
                                         "f"(__uint_as_float(v[4 * j4 + 2])), "f"(__uint_as_float(v[4 * j4 + 3]))
                                         : "memory");
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 256);
    }
}

// ------------------------------------------------------------------------------------------
// The same weight gradient with vertical halo reuse: one unit = one window chunk j, a stage = the (4 + k - 1) x 16 halo of
// that chunk around a 4 x 16 pixel block plus the dY block; the k kernel rows are k shifted views of the halo, stacked two
// by two in the 128 MMA rows (LBO = one image row of the tile), each pair with its own TMEM accumulator.  (4 + k - 1) /
// (4 k) of the X traffic of wgradw_tc_kernel and one dY load / barrier round trip for all k rows.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TC_THREADS, 1)
wgradh_tc_kernel(const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapDY,
                 float* __restrict__ dw, const TcWgradWArgs a) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ygroups = a.Cout / 16;
    const int k = a.k, halo_h = a.Hk + k - 1;
    const uint32_t a_bytes = (uint32_t)(halo_h * a.Wk) * 128u;
    const uint32_t stage_bytes = a_bytes + (uint32_t)ygroups * 2048u;
    const int S = a.stages;
    const uint32_t bar0 = smem0 + S * stage_bytes;
    auto full = [&](int s) { return bar0 + 8u * s; };
    auto empty = [&](int s) { return bar0 + 8u * (S + s); };
    const uint32_t tfull = bar0 + 8u * (2 * S);
    const uint32_t tmem_slot = bar0 + 8u * (2 * S + 1);
    volatile uint32_t* tmem_slot_ptr = (volatile uint32_t*)(smem_raw + (tmem_slot - smem_u32(smem_raw)));
    const int npair = (k + 1) / 2;
    const uint32_t acc_cols = (uint32_t)a.acc_pitch;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&mapX);
        tma_prefetch_desc(&mapDY);
        for (int s = 0; s < S; ++s) { mbar_init(full(s), 1); mbar_init(empty(s), 1); }
        mbar_init(tfull, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, (uint32_t)a.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    griddep_launch();
    griddep_wait();

    const int split = blockIdx.x % a.splits, j = blockIdx.x / a.splits;       // unit = window chunk j
    const int total_chunks = a.nb * a.chunks_per_img;
    const int per = (total_chunks + a.splits - 1) / a.splits;
    const int q_begin = split * per;
    const int q_end = min(total_chunks, q_begin + per);
    const int nq = q_end - q_begin;

    if (warp == 0) {        // uniform control flow for the whole warp, one elected lane issues (see elect_one)
        int s = 0; uint32_t ph = 0;
        for (int q = q_begin; q < q_end; ++q) {
            const int img = q / a.chunks_per_img, r = q % a.chunks_per_img;
            const int w0 = (r % a.chunks_w) * a.Wk, h0 = (r / a.chunks_w) * a.Hk;
            mbar_wait(empty(s), ph ^ 1u);
            const uint32_t sa = smem0 + s * stage_bytes;
            if (elect_one()) {
                mbar_expect_tx(full(s), stage_bytes);
                tma_load_5d(sa, &mapX, full(s), 64 * j, w0, 0, h0 - a.pt, a.n0 + img);
                tma_load_5d(sa + a_bytes, &mapDY, full(s), 0, w0, h0, 0, a.y_n0 + img);
            }
            __syncwarp();
            if (++s == S) { s = 0; ph ^= 1u; }
        }
    } else if (warp == 1) {
        if (nq > 0) {
            int s = 0; uint32_t ph = 0;
            for (int i = 0; i < nq; ++i) {
                const int q = q_begin + i;
                const int w0 = ((q % a.chunks_per_img) % a.chunks_w) * a.Wk;
                mbar_wait(full(s), ph);
                const uint32_t sa = smem0 + s * stage_bytes;
                win_fix(sa, a.Wk, halo_h, w0, a.W, a.C, k, a.pl, 64 * j, lane);
                __syncwarp();
                tc_fence_after();
                if (elect_one()) {
                    for (int pi = 0; pi < npair; ++pi) {
                        // rows 0-63 = kernel row 2*pi, rows 64-127 = kernel row 2*pi+1: the same halo one image row further down
                        const uint64_t adesc = make_smem_desc(sa + (uint32_t)(2 * pi * a.Wk) * 128u, (uint32_t)a.Wk * 128u, 1024);
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk)
                            umma_bf16(tmem_base + (uint32_t)pi * acc_cols, adesc + (uint64_t)(kk * 128),
                                      make_smem_desc32_mn(sa + a_bytes + (uint32_t)kk * 512u, 2048), a.idesc, (uint32_t)((i | kk) != 0));
                    }
                    umma_commit(empty(s));
                    if (i == nq - 1) umma_commit(tfull);
                }
                __syncwarp();
                if (++s == S) { s = 0; ph ^= 1u; }
            }
        }
    } else if (nq > 0) {
        const int q = warp & 3;
        mbar_wait(tfull, 0);
        tc_fence_after();
        const int row = q * 32 + lane, half = row >> 6, i = row & 63;
        const int e = 64 * j + i;
        const int kw = e / a.C, ci = e - kw * a.C;
        const bool live_e = e < k * a.C && ci < a.Creal;
        for (int pi = 0; pi < npair; ++pi) {
            const int kh = 2 * pi + half;
            const bool live = live_e && kh < k;
            float* dst = dw + ((size_t)(kh * k + kw) * a.Creal + ci) * a.Cout;
            const uint32_t taddr = tmem_base + (uint32_t)pi * acc_cols + ((uint32_t)(q * 32) << 16);
            for (int c0 = 0; c0 < a.Cout; c0 += 32) {
                uint32_t v[32];
                tmem_ld32(taddr + (uint32_t)c0, v);
                const int cols = min(32, a.Cout - c0);
                if (live) {
#pragma unroll
                    for (int j4 = 0; j4 < 8; ++j4)
                        if (j4 * 4 < cols)
                            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + c0 + 4 * j4),
                                         "f"(__uint_as_float(v[4 * j4])), "f"(__uint_as_float(v[4 * j4 + 1])),
                                         "f"(__uint_as_float(v[4 * j4 + 2])), "f"(__uint_as_float(v[4 * j4 + 3]))
                                         : "memory");
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, (uint32_t)a.tmem_cols);
    }
}

// packed window weights (see tc_pack_win in conv_tc.h); all window layers of a net in ONE launch: a block finds its
// job in the prefix table and packs a grid-stride share of it
__global__ void pack_win_multi_kernel(const __grid_constant__ TcPackWinJobs jobs) {
    int jb = 0;
    while (jb + 1 < jobs.n && (int)blockIdx.x >= jobs.block0[jb + 1]) ++jb;
    const float* __restrict__ w = jobs.w[jb];
    bf16* __restrict__ wf = jobs.wf[jb];
    const int k = jobs.k[jb], C = jobs.C[jb], Creal = jobs.Creal[jb], n_rows = jobs.n_rows[jb], npad = jobs.npad[jb];
    const int Cin_w = jobs.cin_w[jb], Cout_w = jobs.cout_w[jb], flip = jobs.flip[jb];
    const int nch = (k * C + 63) / 64;
    const size_t total = (size_t)k * nch * npad * 64;
    const int b = blockIdx.x - jobs.block0[jb], nb = jobs.block0[jb + 1] - jobs.block0[jb];
    for (size_t idx = b * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)nb * blockDim.x) {
        const int i = (int)(idx & 63);
        size_t r = idx >> 6;
        const int n = (int)(r % npad); r /= npad;
        // row-major steps (kh, j) for conv_tc_kernel<WIN> / wgradw, chunk-major (j, kh) for convw_tc_kernel (flip bit 1)
        const int j = (flip & 2) ? (int)(r / k) : (int)(r % nch);
        const int s = (flip & 2) ? (int)(r % k) : (int)(r / nch);
        const int e = 64 * j + i, q = e / C, c = e - q * C;
        float v = 0.f;
        if (e < k * C && c < Creal && n < n_rows) {
            // forward: rows = output channels, window element = (kw, ci);  flip: rows = input channels, element = (kw', co)
            const bool fl = (flip & 1) != 0;
            const int kh = fl ? k - 1 - s : s, kw = fl ? k - 1 - q : q;
            const int ci = fl ? n : c, co = fl ? c : n;
            v = w[(((size_t)kh * k + kw) * Cin_w + ci) * Cout_w + co];
        }
        wf[idx] = __float2bfloat16(v);
    }
}

int tc_pack_win_flush(TcPackWinJobs& jobs, cudaStream_t st) {
    if (jobs.n == 0) return CG_OK;
    int total = 0;
    for (int j = 0; j < jobs.n; ++j) {
        const int nch = (jobs.k[j] * jobs.C[j] + 63) / 64;
        const size_t n = (size_t)jobs.k[j] * nch * jobs.npad[j] * 64;
        int blocks = (int)((n + 1023) / 1024);
        if (blocks > 148) blocks = 148;
        jobs.block0[j] = total;
        total += blocks;
    }
    jobs.block0[jobs.n] = total;
    pack_win_multi_kernel<<<total, 256, 0, st>>>(jobs);
    CG_LAUNCH_CHECK();
    jobs.n = 0;
    return CG_OK;
}

int tc_pack_win(TcPackWinJobs& jobs, const float* w, bf16* wf, int k, int C, int Creal, int n_rows, int npad, int Cin_w,
                int Cout_w, int flip, cudaStream_t st) {
    const int j = jobs.n++;
    jobs.w[j] = w; jobs.wf[j] = wf; jobs.k[j] = k; jobs.C[j] = C; jobs.Creal[j] = Creal; jobs.n_rows[j] = n_rows;
    jobs.npad[j] = npad; jobs.cin_w[j] = Cin_w; jobs.cout_w[j] = Cout_w; jobs.flip[j] = flip;
    if (jobs.n == TC_PACK_MAX) return tc_pack_win_flush(jobs, st);
    return CG_OK;
}

// ------------------------------------------------------------------------------------------
// weight packing: float32 HWIO -> bf16 [tap][Cout][Cin] (forward B operand) and bf16 [tap][Cin][Cout]
// (data-gradient B operand; same order as HWIO)
// ------------------------------------------------------------------------------------------
__global__ void pack_weights_kernel(const float* __restrict__ w, bf16* __restrict__ wf, bf16* __restrict__ wd, int taps,
                                    int Cin, int Cout) {
    const size_t n = (size_t)taps * Cin * Cout;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int co = (int)(i % Cout);
        const size_t r = i / Cout;
        const int ci = (int)(r % Cin);
        const int tap = (int)(r / Cin);
        const bf16 v = __float2bfloat16(w[i]);
        wd[i] = v;
        wf[((size_t)tap * Cout + co) * Cin + ci] = v;
    }
}

// all regular tensor-core layers of a net in ONE launch (the per-layer launches were ~9 us each, 50 per step): a block
// finds its layer in the prefix table of the job list and packs a grid-stride share of it
__global__ void pack_weights_multi_kernel(const __grid_constant__ TcPackJobs jobs) {
    int j = 0;
    while (j + 1 < jobs.n && (int)blockIdx.x >= jobs.block0[j + 1]) ++j;
    const float* __restrict__ w = jobs.w[j];
    bf16* __restrict__ wf = jobs.wf[j];
    bf16* __restrict__ wd = jobs.wd[j];
    const int Cin = jobs.cin[j], Cout = jobs.cout[j];
    const size_t n = (size_t)jobs.taps[j] * Cin * Cout;
    const int b = blockIdx.x - jobs.block0[j], nb = jobs.block0[j + 1] - jobs.block0[j];
    for (size_t i = b * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)nb * blockDim.x) {
        const int co = (int)(i % Cout);
        const size_t r = i / Cout;
        const int ci = (int)(r % Cin);
        const int tap = (int)(r / Cin);
        const bf16 v = __float2bfloat16(w[i]);
        wd[i] = v;
        wf[((size_t)tap * Cout + co) * Cin + ci] = v;
    }
}

int tc_pack_weights_multi(TcPackJobs& jobs, cudaStream_t st) {
    if (jobs.n == 0) return CG_OK;
    int total = 0;
    for (int j = 0; j < jobs.n; ++j) {
        const size_t n = (size_t)jobs.taps[j] * jobs.cin[j] * jobs.cout[j];
        int blocks = (int)((n + 255) / 256);
        if (blocks > 148 * 2) blocks = 148 * 2;
        jobs.block0[j] = total;
        total += blocks;
    }
    jobs.block0[jobs.n] = total;
    pack_weights_multi_kernel<<<total, 256, 0, st>>>(jobs);
    CG_LAUNCH_CHECK();
    jobs.n = 0;
    return CG_OK;
}

int tc_pack_weights(const float* w, bf16* wf, bf16* wd, int taps, int Cin, int Cout, cudaStream_t st) {
    size_t n = (size_t)taps * Cin * Cout;
    int blocks = (int)((n + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    pack_weights_kernel<<<blocks, 256, 0, st>>>(w, wf, wd, taps, Cin, Cout);
    CG_LAUNCH_CHECK();
    return CG_OK;
}

// ------------------------------------------------------------------------------------------
// host side: tensor maps and launches
// ------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) == cudaSuccess &&
            qr == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

int tc_make_map_act(CUtensorMap* map, const void* base, int C, int W, int H, int N, int parity, int box_w, int box_h) {
    EncodeTiledFn enc = get_encode();
    if (!enc) { cg_set_error("cuTensorMapEncodeTiled is not available from the driver"); return CG_ERR_CUDA; }
    cuuint64_t dims[5], strides[4];
    if (!parity) {
        dims[0] = C; dims[1] = W; dims[2] = 1; dims[3] = H; dims[4] = N;
        strides[0] = (cuuint64_t)C * 2; strides[1] = (cuuint64_t)W * C * 2; strides[2] = (cuuint64_t)W * C * 2;
        strides[3] = (cuuint64_t)H * W * C * 2;
    } else {
        if ((W | H) & 1) { cg_set_error("parity view needs even H, W (got %dx%d)", H, W); return CG_ERR_INVALID; }
        dims[0] = 2 * (cuuint64_t)C; dims[1] = W / 2; dims[2] = 2; dims[3] = H / 2; dims[4] = N;
        strides[0] = (cuuint64_t)2 * C * 2; strides[1] = (cuuint64_t)W * C * 2; strides[2] = (cuuint64_t)2 * W * C * 2;
        strides[3] = (cuuint64_t)H * W * C * 2;
    }
    cuuint32_t box[5] = {64, (cuuint32_t)box_w, 1, (cuuint32_t)box_h, 1};
    cuuint32_t es[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(base), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        cg_set_error("cuTensorMapEncodeTiled(act C=%d W=%d H=%d N=%d parity=%d box %dx%d) failed: %d", C, W, H, N, parity,
                     box_w, box_h, (int)r);
        return CG_ERR_CUDA;
    }
    return CG_OK;
}

int tc_make_map_act_grouped(CUtensorMap* map, const void* base, int C, int W, int H, int N, int box_w, int box_h, int groups) {
    EncodeTiledFn enc = get_encode();
    if (!enc) { cg_set_error("cuTensorMapEncodeTiled is not available from the driver"); return CG_ERR_CUDA; }
    // the group dimension sits OUTSIDE (w, h) so that the box lands as [group][h][w][64]
    cuuint64_t dims[5] = {64, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)(C / 64), (cuuint64_t)N};
    cuuint64_t strides[4] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, 128, (cuuint64_t)H * W * C * 2};
    cuuint32_t box[5] = {64, (cuuint32_t)box_w, (cuuint32_t)box_h, (cuuint32_t)groups, 1};
    cuuint32_t es[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(base), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        cg_set_error("cuTensorMapEncodeTiled(grouped C=%d W=%d H=%d N=%d box %dx%d x%d) failed: %d", C, W, H, N, box_w, box_h,
                     groups, (int)r);
        return CG_ERR_CUDA;
    }
    return CG_OK;
}

int tc_make_map_act16(CUtensorMap* map, const void* base, int C, int W, int H, int N, int box_w, int box_h, int groups) {
    EncodeTiledFn enc = get_encode();
    if (!enc) { cg_set_error("cuTensorMapEncodeTiled is not available from the driver"); return CG_ERR_CUDA; }
    cuuint64_t dims[5] = {16, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)(C / 16), (cuuint64_t)N};
    cuuint64_t strides[4] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, 32, (cuuint64_t)H * W * C * 2};
    cuuint32_t box[5] = {16, (cuuint32_t)box_w, (cuuint32_t)box_h, (cuuint32_t)groups, 1};
    cuuint32_t es[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(base), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        cg_set_error("cuTensorMapEncodeTiled(act16 C=%d W=%d H=%d N=%d box %dx%d x%d) failed: %d", C, W, H, N, box_w, box_h,
                     groups, (int)r);
        return CG_ERR_CUDA;
    }
    return CG_OK;
}

int tc_make_map_w16(CUtensorMap* map, const void* base, int cols, int rows, int box_rows, int groups) {
    EncodeTiledFn enc = get_encode();
    if (!enc) { cg_set_error("cuTensorMapEncodeTiled is not available from the driver"); return CG_ERR_CUDA; }
    cuuint64_t dims[3] = {16, (cuuint64_t)rows, (cuuint64_t)(cols / 16)};
    cuuint64_t strides[2] = {(cuuint64_t)cols * 2, 32};
    cuuint32_t box[3] = {16, (cuuint32_t)box_rows, (cuuint32_t)groups};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        cg_set_error("cuTensorMapEncodeTiled(w16 cols=%d rows=%d box %d x%d) failed: %d", cols, rows, box_rows, groups, (int)r);
        return CG_ERR_CUDA;
    }
    return CG_OK;
}

int tc_make_map_2d(CUtensorMap* map, const void* base, int cols, int rows, int box_rows) {
    EncodeTiledFn enc = get_encode();
    if (!enc) { cg_set_error("cuTensorMapEncodeTiled is not available from the driver"); return CG_ERR_CUDA; }
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        cg_set_error("cuTensorMapEncodeTiled(2d cols=%d rows=%d box %d) failed: %d", cols, rows, box_rows, (int)r);
        return CG_ERR_CUDA;
    }
    return CG_OK;
}

static int g_num_sms = 0;
static int num_sms() {
    if (!g_num_sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
        if (g_num_sms <= 0) g_num_sms = 148;
    }
    return g_num_sms;
}

int tc_conv_stages(int bn) {
    int stage = A_TILE_BYTES + bn * 128;
    int s = (227 * 1024 - 3072) / stage;
    return s > 6 ? 6 : s;
}

static int g_use_2cta = -1;
int tc_conv_launch(const CUtensorMap* mapA, const CUtensorMap* mapB, const CUtensorMap* mapB2, bf16* out, const float* bias,
                   TcConvArgs a, double flops, cudaStream_t st) {
    // measured on B200 (profiles/r01_tc_kernels.md): the pair kernel lowers L2 traffic (lts 49 % -> 37 %) but not the
    // duration (119 us vs 117 us at the C3 trunk shape), so the 1-CTA kernel stays the default
    if (g_use_2cta < 0) { const char* e = getenv("CG_ENABLE_2CTA"); g_use_2cta = (e && e[0] == '1') ? 1 : 0; }
    static const int dbg = [] { const char* e = getenv("CG_TC_DBG"); return e ? atoi(e) : 0; }();      // measurement only
    a.dbg = dbg;
    if (g_use_2cta && mapB2 && (!a.stats || (dbg & 1)) && !a.bk16 && a.bn >= 32 && ((a.nb * a.tiles_per_img) % 2 == 0)) {
        const int stage = A_TILE_BYTES + (a.bn / 2) * 128;
        int s2 = (227 * 1024 - 2048) / stage;
        a.stages = s2 > 8 ? 8 : s2;
        a.idesc = make_idesc(256, a.bn, 0, 0);
        const size_t smem = (size_t)a.stages * stage + 1024 + 256;
        static std::atomic<unsigned long long> attr2{0};
        if (cg_first_on_device(attr2)) {
            CG_CUDA(cudaFuncSetAttribute(conv_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        }
        const int total_pairs = (a.nb * a.tiles_per_img / 2) * a.n_blocks_n;
        int clusters = num_sms() / 2;
        if (clusters > total_pairs) clusters = total_pairs;
        int pi = prof_begin(st);
        conv_tc2_kernel<<<2 * clusters, TC_THREADS, smem, st>>>(*mapA, *mapB2, out, bias, a);
        prof_end(pi, st, flops, prof_key(2, a.n_taps, a.cchunks, a.bn, a.tiles_per_img, a.nb));
        CG_LAUNCH_CHECK();
        return CG_OK;
    }
    const int stage_b = a.bk16 ? (a.tps > 1 ? a.tps * a.cin16 : a.groups) * (4096 + a.bn * 32) : (A_TILE_BYTES + a.bn * 128);
    // epilogue variant (see conv_tc_kernel): two epilogue groups for the short main loops; CG_EPI1=0 switches them off (A/B),
    // CG_EPI1_KS moves the K-step threshold
    static const int epi1_ks = [] { const char* e = getenv("CG_EPI1"); if (e && e[0] == '0') return -1;
                                    const char* k = getenv("CG_EPI1_KS"); return k ? atoi(k) : 4; }();
    const int ksteps = a.n_taps * a.cchunks;
    const bool plain = !a.bk16 && a.win_C == 0;
    int epi = 0;
    if (plain && ksteps <= epi1_ks) epi = 1;
    const int scratch_w = 32 * 17 * 4;                       // per epilogue warp
    const bool dual = a.bn <= 128 && 2 * stage_b + 1024 + 256 + (epi == 1 ? 8 : 4) * scratch_w <= 110 * 1024;
    const int epi_warps = (dual ? 4 : EPI_WARPS) * (epi == 1 ? 2 : 1);
    const int scratch = epi_warps * scratch_w;
    {
        int sN = ((dual ? 110 : 227) * 1024 - 1024 - 256 - scratch) / stage_b;
        a.stages = sN > 8 ? 8 : sN;
    }
    a.idesc = make_idesc(128, a.bn, 0, 0);
    const size_t smem = (size_t)a.stages * stage_b + 1024 + 256 + scratch;
    static std::atomic<unsigned long long> attr_set{0};
    if (cg_first_on_device(attr_set)) {
        CG_CUDA(cudaFuncSetAttribute(conv_tc_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        CG_CUDA(cudaFuncSetAttribute(conv_tc_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        CG_CUDA(cudaFuncSetAttribute(conv_tc_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024));
        CG_CUDA(cudaFuncSetAttribute(conv_tc_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024));
        CG_CUDA(cudaFuncSetAttribute(conv_tc_kernel<false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        CG_CUDA(cudaFuncSetAttribute(conv_tc_kernel<false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024));
        CG_CUDA(cudaFuncSetAttribute(conv_tc_kernel<false, false, false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        CG_CUDA(cudaFuncSetAttribute(conv_tc_kernel<false, true, false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024));
    }
    const int total = a.nb * a.tiles_per_img * a.n_blocks_n;
    const int slots = (dual ? 2 : 1) * num_sms();
    int grid = total < slots ? total : slots;
    const dim3 threads((2 + epi_warps) * 32);
    int pi = prof_begin(st);
    if (a.win_C > 0) {
        if (dual) launch_pdl(conv_tc_kernel<false, true, true>, dim3(grid), threads, smem, st, *mapA, *mapB, out, bias, a);
        else launch_pdl(conv_tc_kernel<false, false, true>, dim3(grid), threads, smem, st, *mapA, *mapB, out, bias, a);
    } else if (dual) {
        if (a.bk16) launch_pdl(conv_tc_kernel<true, true>, dim3(grid), threads, smem, st, *mapA, *mapB, out, bias, a);
        else if (epi == 1) launch_pdl(conv_tc_kernel<false, true, false, 1>, dim3(grid), threads, smem, st, *mapA, *mapB, out, bias, a);
        else launch_pdl(conv_tc_kernel<false, true>, dim3(grid), threads, smem, st, *mapA, *mapB, out, bias, a);
    } else {
        if (a.bk16) launch_pdl(conv_tc_kernel<true, false>, dim3(grid), threads, smem, st, *mapA, *mapB, out, bias, a);
        else if (epi == 1) launch_pdl(conv_tc_kernel<false, false, false, 1>, dim3(grid), threads, smem, st, *mapA, *mapB, out, bias, a);
        else launch_pdl(conv_tc_kernel<false, false>, dim3(grid), threads, smem, st, *mapA, *mapB, out, bias, a);
    }
    prof_end(pi, st, flops, prof_key(a.win_C > 0 ? 5 : 1, a.n_taps, a.cchunks, a.bn, a.tiles_per_img, a.nb));
    CG_LAUNCH_CHECK();
    return CG_OK;
}

int tc_wgrad_launch(const CUtensorMap* mapX, const CUtensorMap* mapDY, float* dw, TcWgradArgs a, double flops,
                    cudaStream_t st) {
    const int stage = (2 + a.bn / 64) * 8192;
    // N <= 128: a K step is bound by the issue chain (mbarrier wait -> 4 MMAs -> commit, ~600 cycles), not by the tensor
    // pipe, so two CTAs share an SM (<= 110 KB of shared memory and 256 TMEM columns each) and overlap their chains
    const int per_sm = a.bn <= 128 ? 2 : 1;
    int s = ((per_sm == 2 ? 110 : 227) * 1024 - 2048) / stage;
    a.stages = s > 6 ? 6 : s;
    a.idesc = make_idesc(128, a.bn, 1, 1);
    const int units = a.n_taps * a.a_blocks * a.b_blocks;
    const int total_chunks = a.nb * a.chunks_per_img;
    int splits = per_sm * num_sms() / units;
    if (splits < 1) splits = 1;
    if (splits > total_chunks) splits = total_chunks;
    a.splits = splits;
    const size_t smem = (size_t)a.stages * stage + 1024 + 256;
    static std::atomic<unsigned long long> attr_set{0};
    if (cg_first_on_device(attr_set)) {
        CG_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    }
    int pi = prof_begin(st);
    launch_pdl(wgrad_tc_kernel, dim3(units * splits), dim3(TC_THREADS), smem, st, *mapX, *mapDY, dw, a);
    prof_end(pi, st, flops, prof_key(3, a.n_taps, a.a_blocks * a.b_blocks, a.bn, a.chunks_per_img, a.nb));
    CG_LAUNCH_CHECK();
    return CG_OK;
}

int tc_wgrad16_launch(const CUtensorMap* mapX, const CUtensorMap* mapDY, float* dw, TcWgrad16Args a, double flops,
                      cudaStream_t st) {
    const int stage = (8 + a.Cout / 16) * 2048;
    const int per_sm = a.Cout <= 128 ? 2 : 1;              // two resident CTAs overlap their issue chains (see tc_wgrad_launch)
    int s = ((per_sm == 2 ? 110 : 227) * 1024 - 2048) / stage;
    a.stages = s > 8 ? 8 : s;
    a.idesc = make_idesc(128, a.Cout, 1, 1);
    a.m_blocks = (a.n_taps * a.Cin + 127) / 128;
    const int total_chunks = a.nb * a.chunks_per_img;
    int splits = (2 * per_sm * num_sms()) / a.m_blocks;     // ~2 waves of CTAs: the units are short
    if (splits < 1) splits = 1;
    if (splits > total_chunks) splits = total_chunks;
    a.splits = splits;
    const size_t smem = (size_t)a.stages * stage + 1024 + 256;
    static std::atomic<unsigned long long> attr_set{0};
    if (cg_first_on_device(attr_set)) {
        CG_CUDA(cudaFuncSetAttribute(wgrad16_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    }
    int pi = prof_begin(st);
    launch_pdl(wgrad16_tc_kernel, dim3(a.m_blocks * splits), dim3(TC_THREADS), smem, st, *mapX, *mapDY, dw, a);
    prof_end(pi, st, flops, prof_key(4, a.n_taps, a.m_blocks, a.Cout, a.chunks_per_img, a.nb));
    CG_LAUNCH_CHECK();
    return CG_OK;
}

int tc_make_map_win(CUtensorMap* map, const void* x, int C, int k, int pl, int W, int H, int N, int box_w, int box_h) {
    EncodeTiledFn enc = get_encode();
    if (!enc) { cg_set_error("cuTensorMapEncodeTiled is not available from the driver"); return CG_ERR_CUDA; }
    if (C % 8) { cg_set_error("window view needs C %% 8 == 0 (got %d)", C); return CG_ERR_INVALID; }
    const char* base = (const char*)x - (size_t)pl * C * 2;            // coordinate p of dim 1 = window starting at pixel p - pl
    cuuint64_t dims[5] = {(cuuint64_t)k * C, (cuuint64_t)W, 1, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t strides[4] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
    cuuint32_t box[5] = {64, (cuuint32_t)box_w, 1, (cuuint32_t)box_h, 1};
    cuuint32_t es[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<char*>(base), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        cg_set_error("cuTensorMapEncodeTiled(window C=%d k=%d W=%d H=%d N=%d box %dx%d) failed: %d", C, k, W, H, N, box_w, box_h, (int)r);
        return CG_ERR_CUDA;
    }
    return CG_OK;
}

int tc_wgradw_launch(const CUtensorMap* mapX, const CUtensorMap* mapDY, float* dw, TcWgradWArgs a, double flops, cudaStream_t st) {
    static const int force_halves = [] { const char* e = getenv("CG_WGRADW_HALVES"); return e ? atoi(e) : 0; }();    // A/B hook
    a.halves = (a.Cout <= 128 && a.steps >= 4) ? 4 : 2;
    if (force_halves == 2) a.halves = 2;
    const int stage = a.halves * 8192 + (a.Cout / 16) * 2048;
    const int per_sm = stage * 3 + 2048 <= 110 * 1024 ? 2 : 1;       // two resident CTAs overlap their issue chains (see tc_wgrad_launch)
    int s = ((per_sm == 2 ? 110 : 227) * 1024 - 2048) / stage;
    a.stages = s > 8 ? 8 : s;
    a.idesc = make_idesc(128, a.Cout, 1, 1);
    a.units = (a.steps + a.halves - 1) / a.halves;
    const int total_chunks = a.nb * a.chunks_per_img;
    int splits = (2 * per_sm * num_sms()) / a.units;
    if (splits < 1) splits = 1;
    if (splits > total_chunks) splits = total_chunks;
    a.splits = splits;
    const size_t smem = (size_t)a.stages * stage + 1024 + 256;
    static std::atomic<unsigned long long> attr_set{0};
    if (cg_first_on_device(attr_set))
        CG_CUDA(cudaFuncSetAttribute(wgradw_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    int pi = prof_begin(st);
    launch_pdl(wgradw_tc_kernel, dim3(a.units * splits), dim3(TC_THREADS), smem, st, *mapX, *mapDY, dw, a);
    prof_end(pi, st, flops, prof_key(6, a.steps, a.units, a.Cout, a.chunks_per_img, a.nb));
    CG_LAUNCH_CHECK();
    return CG_OK;
}

// window conv with vertical halo reuse: a.Wb x a.Hb pixel boxes, a.cchunks window chunks, a.win_* set; the same packed weights as
// conv_tc_kernel<WIN> ([kh][j][bn][64]).  Returns CG_ERR_INVALID when fewer than two stages fit (the caller then uses the per-row form).
int tc_convw_stages(int bn, int k, int Wb, int Hb, bool* dual) {
    const int stage = (Hb + k - 1) * Wb * 128 + k * bn * 128;
    *dual = bn <= 128 && 2 * stage + 1024 + 256 + EPI_SCRATCH / 2 <= 110 * 1024;
    const int budget = (*dual ? 110 : 227) * 1024 - 1024 - 256 - (*dual ? EPI_SCRATCH / 2 : EPI_SCRATCH);
    int s = budget / stage;
    return s > 6 ? 6 : s;
}

int tc_convw_launch(const CUtensorMap* mapA, const CUtensorMap* mapB, bf16* out, const float* bias, TcConvArgs a, double flops,
                    cudaStream_t st) {
    bool dual = false;
    a.stages = tc_convw_stages(a.bn, a.win_k, a.Wb, a.Hb, &dual);
    if (a.stages < 2) { cg_set_error("window conv: a stage of %d bytes does not fit twice", (a.Hb + a.win_k - 1) * a.Wb * 128 + a.win_k * a.bn * 128); return CG_ERR_INVALID; }
    a.idesc = make_idesc(128, a.bn, 0, 0);
    const int stage = (a.Hb + a.win_k - 1) * a.Wb * 128 + a.win_k * a.bn * 128;
    const size_t smem = (size_t)a.stages * stage + 1024 + 256 + (dual ? EPI_SCRATCH / 2 : EPI_SCRATCH);
    static std::atomic<unsigned long long> attr_set{0};
    if (cg_first_on_device(attr_set)) {
        CG_CUDA(cudaFuncSetAttribute(convw_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        CG_CUDA(cudaFuncSetAttribute(convw_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024));
    }
    const int total = a.nb * a.tiles_per_img;
    const int slots = (dual ? 2 : 1) * num_sms();
    const int grid = total < slots ? total : slots;
    int pi = prof_begin(st);
    if (dual) launch_pdl(convw_tc_kernel<true>, dim3(grid), dim3(TC_THREADS), smem, st, *mapA, *mapB, out, bias, a);
    else launch_pdl(convw_tc_kernel<false>, dim3(grid), dim3(CONV_THREADS), smem, st, *mapA, *mapB, out, bias, a);
    prof_end(pi, st, flops, prof_key(7, a.win_k * a.cchunks, a.cchunks, a.bn, a.tiles_per_img, a.nb));
    CG_LAUNCH_CHECK();
    return CG_OK;
}

// halo form of the window weight gradient (wgradh_tc_kernel): a.Wk x a.Hk = 16 x 4 pixel blocks, a.pt set; returns
// CG_ERR_INVALID when the accumulators do not fit TMEM or fewer than two stages fit
int tc_wgradh_ok(int k, int Cout) {
    const int npair = (k + 1) / 2, pitch = Cout <= 32 ? 32 : Cout <= 64 ? 64 : Cout <= 128 ? 128 : 256;
    return npair * pitch <= 512;
}

int tc_wgradh_launch(const CUtensorMap* mapX, const CUtensorMap* mapDY, float* dw, TcWgradWArgs a, double flops, cudaStream_t st) {
    const int npair = (a.k + 1) / 2;
    a.acc_pitch = a.Cout <= 32 ? 32 : a.Cout <= 64 ? 64 : a.Cout <= 128 ? 128 : 256;
    int cols = npair * a.acc_pitch;
    if (cols > 512) { cg_set_error("wgradh: %d accumulators of %d columns exceed TMEM", npair, a.acc_pitch); return CG_ERR_INVALID; }
    int tm = 32;
    while (tm < cols) tm *= 2;
    a.tmem_cols = tm;
    const int stage = (a.Hk + a.k - 1) * a.Wk * 128 + (a.Cout / 16) * 2048;
    const int per_sm = (tm <= 256 && stage * 3 + 2048 <= 110 * 1024) ? 2 : 1;
    int s = ((per_sm == 2 ? 110 : 227) * 1024 - 2048) / stage;
    a.stages = s > 8 ? 8 : s;
    if (a.stages < 2) { cg_set_error("wgradh: stage of %d bytes does not fit twice", stage); return CG_ERR_INVALID; }
    a.idesc = make_idesc(128, a.Cout, 1, 1);
    a.units = a.nch;
    const int total_chunks = a.nb * a.chunks_per_img;
    int splits = (2 * per_sm * num_sms()) / a.units;
    if (splits < 1) splits = 1;
    if (splits > total_chunks) splits = total_chunks;
    a.splits = splits;
    const size_t smem = (size_t)a.stages * stage + 1024 + 256;
    static std::atomic<unsigned long long> attr_set{0};
    if (cg_first_on_device(attr_set))
        CG_CUDA(cudaFuncSetAttribute(wgradh_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    int pi = prof_begin(st);
    launch_pdl(wgradh_tc_kernel, dim3(a.units * splits), dim3(TC_THREADS), smem, st, *mapX, *mapDY, dw, a);
    prof_end(pi, st, flops, prof_key(8, a.steps, a.units, a.Cout, a.chunks_per_img, a.nb));
    CG_LAUNCH_CHECK();
    return CG_OK;
}
