// track_persistent.cuh -- SE3Tracker::trackFrame as ONE persistent cooperative kernel (mode 1).
//
// The whole coarse-to-fine Levenberg-Marquardt loop of Tracking/SE3Tracker.cpp:280-486 runs on the device:
// every CTA evaluates its share of the keyframe pixels with evalPoint() (track.cuh), the CTAs exchange
// EV_NCH partial sums through L2 behind ONE grid-wide barrier per evaluation, and then EVERY CTA redundantly
// (and deterministically: same instructions on the same data) combines the partials in block order, solves
// the damped 6x6 system (LDL^T), applies exp(inc) * T and takes the accept / reject / converge decision of
// :381-446.  No host round trip, no second barrier, no float atomics.  The host sees one launch and one
// 200-byte result per frame.
//
// Why this shape on B200: a 640x480 frame has <= 77k points on the finest tracked level, i.e. ~2 per resident
// thread; one evaluation is a few microseconds of latency, so the reference's structure (3 passes per
// evaluation, one host decision between evaluations, 15-60 evaluations per frame) is launch/latency bound by
// two orders of magnitude.  Grid = one CTA per SM (148), co-residency guaranteed by the cooperative launch.
#pragma once
#include "internal.cuh"
#include "track.cuh"
#include <stdlib.h>
#include <atomic>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

#define TP_THREADS 512
#define TP_MAXGRID_DBG 160
#define TP_WIN_SMEM ((TP_THREADS / 32) * TRK_WIN_W * TRK_WIN_H * 16)    // 16 warps x 13056 B = 208896 B of dynamic smem
#define EX_ROW 48                    // floats per exchange row (192 B): EV_NCH data + tag + pad
#define TP_LOCAL_MAX_PIXELS 1024      // levels up to this many pixels are evaluated redundantly per CTA
#define TP_CLUSTER_MAX_PIXELS 8192    // ... and up to this many redundantly per thread-block cluster (DSMEM exchange)

struct TrackLevelParams {
    const float* kfIdepth;
    const float* kfVar;
    const float* kfColor;
    const float4* frameGrad;
    int w, h;
    float fx, fy, cx, cy, fxi, fyi, cxi, cyi;
};

struct alignas(64) TrackParams {
    CUtensorMap gradMap[LSD_LEVELS]; // TMA descriptors of the tracked frame's gradient levels (64-byte aligned)
    TrackLevelParams lvl[LSD_LEVELS];
    uint8_t* goodMask;               // level-1 mask of the tracked frame
    int maskFresh;                   // 1: initialise the mask to true first (Frame.h:433)
    int maskBytes;
    int W, H;                        // tracker construction size (diverge test uses width>>lvl, SE3Tracker.cpp:324)
    float initRefToFrame[7];         // frameToReference_initialEstimate.inverse().cast<float>(), :306
    lsdgpu_track_settings st;
    EvalConsts C;
    int useAffine;
    float* partials;                 // [2][EV_NCH][gridDim]
    unsigned int* barrier;           // [0] arrival counter (monotonic), [32] released-epoch flag
    int barrierMode;
    unsigned int barrierBase;        // arrivals counted by earlier launches (the counter is never reset)
    int debug;                       // 1: also write the per-CTA cycle table
    unsigned int launchSeq;          // value the kernel stores in TrackState::doneSeq when the result block is complete
    int minLevel;                    // last level of the coarse-to-fine loop (1 for trackFrame, 4 for permaRef tracking)
    int clusterLocalMaxPixels;       // levels up to this size are evaluated per cluster (0: never; needs a cluster launch)
    int useTma;                      // 1: per-warp shared-memory windows loaded by TMA; 0: all taps through L1/L2
    int doPrepare;                   // 1: the last thread also turns the result into the observe parameters of the same frame
    PrepareConsts prep;              //    (Frame::prepareForStereoWith + head of DepthMap::updateKeyframe), no host round trip
    ObserveParams* obsOut;
    int* skipOut;
};

// what the kernel hands back (block 0 writes it)
struct TrackState {
    float refToFrame[7];
    float pointUsage, goodCount, badCount, meanRes, lastResidual;
    float affine_a, affine_b;
    int diverged;
    int numCalcResidualCalls[LSD_LEVELS];
    int numCalcWarpUpdateCalls[LSD_LEVELS];
    int totalEvals;
    volatile unsigned int doneSeq;   // written last (after a system-wide fence): the host polls it instead of a stream sync
    long long cyc[6];                // block-0 cycle breakdown: points, CTA reduce, barrier, combine, serial LM, total
    long long cycBlk[TP_MAXGRID_DBG][6];   // the same per CTA (debug)
};

// ---- TMA / mbarrier PTX (sm_90+; SASS: UTMALDG / SYNCS) ---------------------------------------------------
__device__ __forceinline__ uint32_t smemU32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbarInit(uint64_t* bar, unsigned int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smemU32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbarExpectTx(uint64_t* bar, unsigned int bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smemU32(bar)), "r"(bytes) : "memory");
}
// bounded wait: returns false if the phase did not complete within ~maxPolls probes (never observed; keeps a
// mis-programmed copy from hanging the whole cooperative grid)
__device__ __forceinline__ bool mbarWait(uint64_t* bar, unsigned int parity, int maxPolls = 1 << 16)
{
    for (int it = 0; it < maxPolls; it++) {
        unsigned int done;
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n" : "=r"(done) : "r"(smemU32(bar)), "r"(parity) : "memory");
        if (done) return true;
    }
    return false;
}
__device__ __forceinline__ void tmaLoad2D(void* dst, const CUtensorMap* map, int x, int y, uint64_t* bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smemU32(dst)), "l"(map), "r"(x), "r"(y), "r"(smemU32(bar)) : "memory");
}

// Grid-wide barrier.  Arrivals are counted on `counter[0]` (monotonic: epoch e completes at e * gridDim
// arrivals); the LAST arriver publishes the epoch number in `counter[32]` (a different 128-byte line), which is
// what everybody else polls -- the pollers never touch the line the atomics serialise on.  Thread 0 carries the
// release / acquire for its CTA (the fences are cumulative over the preceding / following __syncthreads).
__device__ __forceinline__ void gridBarrier(unsigned int* counter, unsigned int base, unsigned int epoch, int mode)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int target = base + epoch * gridDim.x;     // wraps consistently (unsigned arithmetic)
        if (mode == 0) {
            // arrival counter + separate release flag (last arriver publishes the epoch)
            __threadfence();
            const unsigned int old = atomicAdd(counter, 1u);
            volatile unsigned int* flag = counter + 32;
            if (old == target - 1u) { __threadfence(); *flag = target; }
            else { while ((int)(*flag - target) < 0) { } }
            __threadfence();
        } else if (mode == 1) {
            // release-add, then acquire-poll the counter itself (one L2 hop less than mode 0)
            unsigned int v;
            asm volatile("atom.add.release.gpu.global.u32 %0, [%1], 1;" : "=r"(v) : "l"(counter) : "memory");
            if (v != target - 1u) {
                do {
                    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
                } while ((int)(v - target) < 0);
            } else {
                asm volatile("fence.acq_rel.gpu;" ::: "memory");
            }
        } else {
            // fire-and-forget release reduction + acquire poll
            asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
            unsigned int v;
            do {
                asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
            } while ((int)(v - target) < 0);
        }
    }
    __syncthreads();
}

struct LMShared {
    float sums[EV_NCH];
    EvalPose pose;                   // pose being evaluated
    int action;
    int lvl;                         // level of the next evaluation
};

enum { ACT_CONTINUE = 0, ACT_LEVEL_DONE = 1, ACT_DIVERGED = 2, ACT_RETRY = 3, ACT_ACCEPTED = 4 };

__device__ __forceinline__ void setEvalPose(EvalPose& P, const lsd::SE3<float>& T, float a, float b)
{
    lsd::quatToMatrix(T.q, P.R);
    P.t[0] = T.t[0]; P.t[1] = T.t[1]; P.t[2] = T.t[2];
    P.a = a; P.b = b;
}

// ---- multi-value warp reduction -------------------------------------------------------------------------
// K values per lane (K a power of two <= 32) are reduced across the warp with K-1 (+ log2(32/K)) shuffles
// instead of 5K: at every step the lanes of a pair exchange HALF of their values.  On return v[0] holds the
// warp total of channel  chan = sum over used masks of (lane & mask ? half : 0)  (for K = 32: chan == lane).
// Fixed exchange pattern => bit-identical results wherever the same data is reduced.
template <int K, typename T>
__device__ __forceinline__ void warpReduceMulti(T (&v)[K], int lane)
{
    int mask = 16;
#pragma unroll
    for (int half = K / 2; half >= 1; half >>= 1, mask >>= 1) {
        const bool upper = (lane & mask) != 0;
#pragma unroll
        for (int i = 0; i < half; i++) {
            T keep = upper ? v[i + half] : v[i];
            T send = upper ? v[i] : v[i + half];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, mask);
        }
    }
#pragma unroll
    for (; mask >= 1; mask >>= 1) v[0] += __shfl_xor_sync(0xffffffffu, v[0], mask);
}
// channel held in v[0] by `lane` after warpReduceMulti<K>
template <int K> __device__ __forceinline__ int warpReduceChannel(int lane)
{
    int c = 0, mask = 16;
#pragma unroll
    for (int half = K / 2; half >= 1; half >>= 1, mask >>= 1)
        if (lane & mask) c += half;
    return c;
}

// all EV_NCH (= 32 + 8) channels of a thread's accumulator -> per-warp totals in sm[warp][*] (39 shuffles)
__device__ __forceinline__ void warpReduceAcc(PointAcc& acc, int lane, float* smRow)
{
    float a32[32], a8[8];
#pragma unroll
    for (int i = 0; i < 32; i++) a32[i] = acc.v[i];
#pragma unroll
    for (int i = 0; i < 8; i++) a8[i] = acc.v[32 + i];
    warpReduceMulti<32>(a32, lane);
    warpReduceMulti<8>(a8, lane);
    smRow[lane] = a32[0];
    if ((lane & 3) == 0) smRow[32 + warpReduceChannel<8>(lane)] = a8[0];
}

#define TP_WARPS (TP_THREADS / 32)
#define TP_MAXGRID 160               // combine code is unrolled for <= 160 CTAs (B200: 148 SMs)

// One evaluation.  Levels with few pixels (local == true) are evaluated REDUNDANTLY by every CTA over the whole
// level -- no grid barrier, no exchange; big levels are split over the grid with one barrier.  On return
// sh.sums holds the EV_NCH totals, bit-identical in every CTA.
struct WarpWindow {
    const float4* win;               // this warp's TRK_WIN_H x TRK_WIN_W window in shared memory
    uint64_t* bar;
    unsigned int parity;             // phase of the next completion to wait for
    int lvl;                         // level the window currently holds (-1: none)
    int ox, oy;                      // window origin in level pixels
    bool valid;
    unsigned int hits;               // low 16 bits: taps served from the window, high 16: taps through L1/L2
    // the thread's first pixel of the current level, reconstructed once (see gridEvaluate)
    int pcLvl, pcMode;
    bool pcIsPoint;
    float pcx, pcy, pcz, pcVar, pcColor;
};

enum { EVAL_GRID = 0, EVAL_CTA = 1, EVAL_CLUSTER = 2 };

// evalMode: EVAL_GRID    the level is split over the whole grid (one grid barrier + L2 exchange),
//           EVAL_CLUSTER every thread-block cluster evaluates the whole level redundantly (small levels: the CTAs of a
//                        cluster exchange their 40 sums through distributed shared memory behind one hardware
//                        cluster barrier, ~10x cheaper than the grid barrier and no L2 round trip),
//           EVAL_CTA     every CTA evaluates the whole level redundantly (tiny levels, no exchange at all).
__device__ __forceinline__ void gridEvaluate(const TrackParams& p, int lvl, int evalMode, LMShared& sh, float (*sm)[EV_NCH],
                                             float (*xrow)[EV_NCH], unsigned int& xphase,
                                             unsigned int& epoch, long long* cyc, WarpWindow& W)
{
    const bool local = evalMode != EVAL_GRID;
    long long t0 = clock64();
    const TrackLevelParams& L = p.lvl[lvl];
    const EvalPose P = sh.pose;
    PointAcc acc;
#pragma unroll
    for (int c = 0; c < EV_NCH; c++) acc.v[c] = 0.f;
    const int w = L.w, h = L.h, n = w * h;
    const float4* fg = L.frameGrad;
    uint8_t* mask = (lvl == SE3TRACKING_MIN_LEVEL) ? p.goodMask : nullptr;
    // 32-pixel chunks are dealt round-robin to the CTAs (chunk c -> CTA c % G, warp (c / G) % TP_WARPS): the
    // semi-dense density varies over the image, a contiguous split would leave CTAs unevenly loaded
    cg::cluster_group cluster = cg::this_cluster();
    const int crank = (int)cluster.block_rank(), csize = (int)cluster.num_blocks();
    int first, stride;
    if (evalMode == EVAL_GRID) { first = ((threadIdx.x >> 5) * gridDim.x + blockIdx.x) * 32 + (threadIdx.x & 31); stride = gridDim.x * TP_THREADS; }
    else if (evalMode == EVAL_CLUSTER) { first = ((threadIdx.x >> 5) * csize + crank) * 32 + (threadIdx.x & 31); stride = csize * TP_THREADS; }
    else { first = threadIdx.x; stride = TP_THREADS; }
    // the loop bound is evaluated on the chunk base so that whole warps stay in the loop (the window set-up below
    // uses full-warp ballots / shuffles); w*h need not be a multiple of 32 (e.g. 40x30 on level 4)
    const int laneId = threadIdx.x & 31;
    for (int base = first - laneId; base < n; base += stride) {
        const int i = base + laneId;
        const int x = i % w, y = i / w;
        bool isPoint = false;
        float px = 0.f, py = 0.f, pz = 0.f, var = 0.f, color = 0.f;
        const bool firstChunk = (base == first - laneId);
        if (firstChunk && W.pcLvl == lvl && W.pcMode == evalMode) {
            // the thread's first pixel of a level never changes between the evaluations of that level: keep the
            // reconstructed point in registers instead of re-reading the keyframe planes (an L2 round trip at the head
            // of every evaluation's dependency chain)
            isPoint = W.pcIsPoint; px = W.pcx; py = W.pcy; pz = W.pcz; var = W.pcVar; color = W.pcColor;
        } else {
            if (i < n && x >= 1 && x < w - 1 && y >= 1 && y < h - 1) {
                const float idepth = __ldg(L.kfIdepth + i);
                var = __ldg(L.kfVar + i);
                if (!(var <= 0 || idepth == 0)) {
                    const float sc = 1.0f / idepth;
                    px = sc * (L.fxi * x + L.cxi); py = sc * (L.fyi * y + L.cyi); pz = sc * 1;
                    color = __ldg(L.kfColor + i);
                    isPoint = true;
                }
            }
            if (firstChunk) {
                W.pcLvl = lvl; W.pcMode = evalMode; W.pcIsPoint = isPoint;
                W.pcx = px; W.pcy = py; W.pcz = pz; W.pcVar = var; W.pcColor = color;
            }
        }
        // First chunk of this warp on a new level: stage the part of the frame's gradient level that the chunk
        // warps into (centred on the chunk under the level's initial pose) in shared memory with ONE TMA box copy.
        // All later evaluations of the level tap the window; taps that leave it fall back to L1/L2.
        if (p.useTma && !local && firstChunk && W.lvl != lvl) {          // warp-uniform condition
            const int lane = threadIdx.x & 31;
            float u = 0.f, v = 0.f;
            bool ok = false;
            if (isPoint) {
                const float Wx = ((P.R[0] * px + P.R[1] * py) + P.R[2] * pz) + P.t[0];
                const float Wy = ((P.R[3] * px + P.R[4] * py) + P.R[5] * pz) + P.t[1];
                const float Wz = ((P.R[6] * px + P.R[7] * py) + P.R[8] * pz) + P.t[2];
                u = (Wx / Wz) * L.fx + L.cx; v = (Wy / Wz) * L.fy + L.cy;
                ok = (u > -1e4f && u < 1e4f && v > -1e4f && v < 1e4f);
            }
            const unsigned int bal = __ballot_sync(0xffffffffu, ok);
            W.lvl = lvl;
            W.valid = bal != 0u;
            if (W.valid) {
                // reference lane: the valid lane closest to the chunk centre
                int src = 16;
                if (!((bal >> 16) & 1u)) src = (bal >> 16) ? (__ffs(bal >> 16) - 1 + 16) : (31 - __clz(bal));
                const float uc = __shfl_sync(0xffffffffu, u, src), vc = __shfl_sync(0xffffffffu, v, src);
                W.ox = (int)floorf(uc) - (src - 16) - (TRK_WIN_W - 32) / 2 - 16;
                W.oy = (int)floorf(vc) - TRK_WIN_H / 2;
                __syncwarp();
                if (lane == 0) {
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic reads of the old window are done
                    mbarExpectTx(W.bar, TRK_WIN_W * TRK_WIN_H * 16);
                    tmaLoad2D((void*)W.win, &p.gradMap[lvl], 4 * W.ox, W.oy, W.bar);
                }
                if (mbarWait(W.bar, W.parity)) W.parity ^= 1u;
                else { W.valid = false; W.lvl = -2; if (lane == 0) atomicAdd(p.barrier + 40, 1u); }   // counted, reported by the host
            }
        }
        if (isPoint) {
            const bool useWin = p.useTma && !local && firstChunk && W.valid && W.lvl == lvl;
            const float4* win = W.win;
            const int ox = W.ox, oy = W.oy;
            unsigned int* hitCtr = &W.hits;
            auto tap = [fg, w, win, ox, oy, useWin, hitCtr](float u, float v, float& o0, float& o1, float& o2) {
                const int ix = (int)u, iy = (int)v;
                const int lx = ix - ox, ly = iy - oy;
                if (useWin && lx >= 0 && lx < TRK_WIN_W - 1 && ly >= 0 && ly < TRK_WIN_H - 1) {
                    (*hitCtr)++;
                    // same weights and summation order as interp43 (globalFuncs.h:63-77), texels from shared memory
                    const float dx = u - ix, dy = v - iy;
                    const float dxdy = dx * dy;
                    const float4* bp = win + ly * TRK_WIN_W + lx;
                    const float4 br = bp[1 + TRK_WIN_W], bl = bp[TRK_WIN_W], tr = bp[1], tl = bp[0];
                    const float w0 = dxdy, w1 = (dy - dxdy), w2 = (dx - dxdy), w3 = (1 - dx - dy + dxdy);
                    o0 = w0 * br.x + w1 * bl.x + w2 * tr.x + w3 * tl.x;
                    o1 = w0 * br.y + w1 * bl.y + w2 * tr.y + w3 * tl.y;
                    o2 = w0 * br.z + w1 * bl.z + w2 * tr.z + w3 * tl.z;
                } else {
                    (*hitCtr) += 0x10000u;
                    interp43(fg, u, v, w, o0, o1, o2);
                }
            };
            int good = evalPoint(px, py, pz, color, var, P, p.C, L.fx, L.fy, L.cx, L.cy, w, h, tap, acc);
            if (mask) mask[i] = (uint8_t)good;
        }
    }
    __syncthreads();
    long long t1 = clock64();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    warpReduceAcc(acc, lane, sm[warp]);
    __syncthreads();
    float ctaSum = 0.f;
    if (threadIdx.x < EV_NCH) {
#pragma unroll
        for (int wi = 0; wi < TP_WARPS; wi++) ctaSum += sm[wi][threadIdx.x];
    }
    if (evalMode == EVAL_CTA) {
        if (threadIdx.x < EV_NCH) sh.sums[threadIdx.x] = ctaSum;
        __syncthreads();
        long long t2 = clock64();
        cyc[0] += t1 - t0; cyc[1] += t2 - t1;
        return;
    }
    if (evalMode == EVAL_CLUSTER) {
        // DSMEM exchange: publish this CTA's row (double-buffered by phase), one cluster barrier, then every CTA sums
        // the rows of all ranks in rank order -> identical totals in every CTA of every cluster
        float* mine = xrow[xphase & 1u];
        if (threadIdx.x < EV_NCH) mine[threadIdx.x] = ctaSum;
        cluster.sync();
        if (threadIdx.x < EV_NCH) {
            float tot = 0.f;
            for (int r = 0; r < csize; r++) tot += cluster.map_shared_rank(mine, r)[threadIdx.x];
            sh.sums[threadIdx.x] = tot;
        }
        xphase++;
        __syncthreads();
        long long t2 = clock64();
        cyc[0] += t1 - t0; cyc[1] += t2 - t1;
        return;
    }
    // ---- exchange: one partial row per CTA ([parity][channel][cta], written through to L2), ONE grid barrier,
    // then every CTA combines all rows in a fixed order (measured alternatives: per-row tag polling without a
    // barrier is slower -- 148 x 148 polling loads saturate L2)
    const unsigned int parity = epoch & 1u;
    float* part = p.partials + (size_t)parity * EV_NCH * gridDim.x;
    if (threadIdx.x < EV_NCH) __stcg(part + (size_t)threadIdx.x * gridDim.x + blockIdx.x, ctaSum);
    epoch++;
    long long t2 = clock64();
    gridBarrier(p.barrier, p.barrierBase, epoch, p.barrierMode);
    long long t3 = clock64();
    // warp wi owns channels wi, wi+TP_WARPS, ... (<= 3 per warp); all loads are issued first, then one
    // multi-value reduction in double
    {
        double ch[4];
        const int nb = (int)gridDim.x;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int c = warp + k * TP_WARPS;
            float v[TP_MAXGRID / 32];
#pragma unroll
            for (int j = 0; j < TP_MAXGRID / 32; j++) {
                const int bb = lane + 32 * j;
                v[j] = (c < EV_NCH && bb < nb) ? __ldcg(part + (size_t)c * nb + bb) : 0.f;
            }
            double sacc = 0.0;
#pragma unroll
            for (int j = 0; j < TP_MAXGRID / 32; j++) sacc += (double)v[j];
            ch[k] = sacc;
        }
        warpReduceMulti<4>(ch, lane);
        if ((lane & 7) == 0) {
            const int c = warp + warpReduceChannel<4>(lane) * TP_WARPS;
            if (c < EV_NCH) sh.sums[c] = (float)ch[0];
        }
    }
    __syncthreads();
    long long t4 = clock64();
    cyc[0] += t1 - t0; cyc[1] += t2 - t1; cyc[2] += t3 - t2; cyc[3] += t4 - t3;
}

// Fully unrolled, register-resident LDL^T (no pivoting) for the damped 6x6 normal equations.  Returns false if
// a pivot is not strictly positive (then the caller falls back to the pivoted routine, hostmath.h).
__device__ __forceinline__ bool ldlt6SolveFast(const float* A, const float* b, float* x)
{
    float L[6][6], D[6];
    bool ok = true;
#pragma unroll
    for (int j = 0; j < 6; j++) {
        float d = A[j * 6 + j];
#pragma unroll
        for (int k = 0; k < j; k++) d -= L[j][k] * L[j][k] * D[k];
        D[j] = d;
        ok = ok && (d > 0.f);
        const float inv = 1.0f / d;
#pragma unroll
        for (int i = j + 1; i < 6; i++) {
            float v = A[i * 6 + j];
#pragma unroll
            for (int k = 0; k < j; k++) v -= L[i][k] * L[j][k] * D[k];
            L[i][j] = v * inv;
        }
    }
    float y[6];
#pragma unroll
    for (int i = 0; i < 6; i++) {
        float s = b[i];
#pragma unroll
        for (int k = 0; k < i; k++) s -= L[i][k] * y[k];
        y[i] = s;
    }
#pragma unroll
    for (int i = 0; i < 6; i++) y[i] = y[i] / D[i];
#pragma unroll
    for (int i = 5; i >= 0; i--) {
        float s = y[i];
#pragma unroll
        for (int k = i + 1; k < 6; k++) s -= L[k][i] * x[k];
        x[i] = s;
    }
    return ok;
}

// exp(inc) * T on the device, latency-trimmed: one sincosf (half angle; sin/cos of the full angle by the
// double-angle identities), reciprocal-sqrt normalisations.  Same algebra as lsd::se3Exp / se3Mul (hostmath.h,
// se3.hpp:406-428, 239-259); results agree to ~1e-7 relative.
__device__ __forceinline__ lsd::SE3<float> se3ExpMulFast(const float* a, const lsd::SE3<float>& T)
{
    const float wx = a[3], wy = a[4], wz = a[5];
    const float theta_sq = wx * wx + wy * wy + wz * wz;
    float imag, real, c1, c2;
    if (theta_sq < 1e-10f) {
        const float theta_po4 = theta_sq * theta_sq;
        imag = 0.5f - (1.0f / 48.0f) * theta_sq + (1.0f / 3840.0f) * theta_po4;
        real = 1.0f - 0.5f * theta_sq + (1.0f / 384.0f) * theta_po4;
        c1 = 0.5f; c2 = 1.0f / 6.0f;                       // V = I + Omega/2 + Omega^2/6 (theta -> 0)
    } else {
        const float inv_theta = rsqrtf(theta_sq);
        const float theta = theta_sq * inv_theta;
        float sh_, ch_;
        sincosf(0.5f * theta, &sh_, &ch_);
        imag = sh_ * inv_theta;
        real = ch_;
        const float inv_sq = inv_theta * inv_theta;
        c1 = 2.0f * sh_ * sh_ * inv_sq;                     // (1 - cos theta) / theta^2
        c2 = (theta - 2.0f * sh_ * ch_) * inv_sq * inv_theta;   // (theta - sin theta) / theta^3
    }
    float q[4] = { imag * wx, imag * wy, imag * wz, real };
    {
        const float n = rsqrtf(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
        q[0] *= n; q[1] *= n; q[2] *= n; q[3] *= n;
    }
    // V * upsilon with V = I + c1*Omega + c2*Omega^2 :  Omega u = w x u,  Omega^2 u = w x (w x u)
    const float ux = a[0], uy = a[1], uz = a[2];
    const float k1x = wy * uz - wz * uy, k1y = wz * ux - wx * uz, k1z = wx * uy - wy * ux;
    const float k2x = wy * k1z - wz * k1y, k2y = wz * k1x - wx * k1z, k2z = wx * k1y - wy * k1x;
    const float tx = ux + c1 * k1x + c2 * k2x, ty = uy + c1 * k1y + c2 * k2y, tz = uz + c1 * k1z + c2 * k2z;
    // (q, t) * T
    lsd::SE3<float> r;
    float rt[3];
    lsd::quatRotate(q, T.t, rt);
    lsd::quatMul(q, T.q, r.q);
    {
        const float n = rsqrtf(r.q[0] * r.q[0] + r.q[1] * r.q[1] + r.q[2] * r.q[2] + r.q[3] * r.q[3]);
        r.q[0] *= n; r.q[1] *= n; r.q[2] *= n; r.q[3] *= n;
    }
    r.t[0] = tx + rt[0]; r.t[1] = ty + rt[1]; r.t[2] = tz + rt[2];
    return r;
}

// LM state of SE3Tracker::trackFrame; lives in shared memory, touched by thread 0 only (keeps it out of the
// register budget of the other threads)
struct LMState {
    lsd::SE3<float> refToFrame, cand;
    float affine_a, affine_b, lastErr, last_residual, LM_lambda;
    float inc[6];
    float lsq[27];                   // accepted normal equations, RAW sums: 21 upper-triangle A + 6 (sum J r w)
    int nRes[LSD_LEVELS], nUpd[LSD_LEVELS];
    int lvl, iteration, phase, incTry, diverged;
    long long dbg[4];                // thread-0 cycles: decision, solve, pose update, (spare)
};
enum { PH_INIT = 0, PH_TRY = 1 };

// affine lighting estimate of the last evaluation, SE3Tracker.cpp:1023-1024
__device__ __forceinline__ void affineFromSums(const float* s, float& a, float& b)
{
    const float sxx = s[CH_SXX], syy = s[CH_SYY], sx = s[CH_SX], sy = s[CH_SY], sw = s[CH_SW];
    a = sqrtf((syy - sy * sy / sw) / (sxx - sx * sx / sw));
    b = (sy - a * sx) / sw;
}

// solve the damped system of the accepted linearisation and set the candidate pose (SE3Tracker.cpp:356-363).
// A and b are used un-normalised: LGS6::finish divides both by num_constraints (LGSX.h:319-325), which cancels
// in A^-1 b (the damping is multiplicative), so the division is skipped on the device.
__device__ __forceinline__ void lmSolveAndPropose(LMState& lm, LMShared& sh)
{
    static const unsigned char ij[21][2] = { {0,0},{0,1},{0,2},{0,3},{0,4},{0,5},{1,1},{1,2},{1,3},{1,4},{1,5},
                                            {2,2},{2,3},{2,4},{2,5},{3,3},{3,4},{3,5},{4,4},{4,5},{5,5} };
    float A[36], b[6], inc[6];
    const long long ta = clock64();
#pragma unroll
    for (int k = 0; k < 21; k++) {
        const float v = lm.lsq[k];
        A[ij[k][0] * 6 + ij[k][1]] = v;
        A[ij[k][1] * 6 + ij[k][0]] = v;
    }
#pragma unroll
    for (int i = 0; i < 6; i++) b[i] = lm.lsq[21 + i];
    const float damp = 1 + lm.LM_lambda;
#pragma unroll
    for (int i = 0; i < 6; i++) A[i * 6 + i] *= damp;
    if (!ldlt6SolveFast(A, b, inc)) lsd::ldlt6Solve(A, b, inc);
#pragma unroll
    for (int i = 0; i < 6; i++) lm.inc[i] = inc[i];
    const long long tb = clock64();
    lm.incTry++;
    lm.cand = se3ExpMulFast(inc, lm.refToFrame);
    setEvalPose(sh.pose, lm.cand, lm.affine_a, lm.affine_b);
    lm.phase = PH_TRY;
    const long long tc = clock64();
    lm.dbg[1] += tb - ta; lm.dbg[2] += tc - tb;
}

__device__ __forceinline__ void lmNextLevel(const TrackParams& p, LMState& lm, LMShared& sh)
{
    lm.lvl--;
    if (lm.lvl < p.minLevel) { sh.action = ACT_LEVEL_DONE; return; }                // all levels done
    lm.phase = PH_INIT;
    setEvalPose(sh.pose, lm.refToFrame, lm.affine_a, lm.affine_b);
    sh.lvl = lm.lvl;
}

// thread 0 after every evaluation: the decisions of SE3Tracker.cpp:324-446.  Written as one straight line with a
// single solve/propose site (the three call sites of the reference's loop structure -- first iteration of a level,
// next iteration after an accept, retry after a reject -- only differ in which normal equations and lambda they use).
__device__ __forceinline__ void lmAdvance(const TrackParams& p, LMState& lm, LMShared& sh)
{
    const float* s = sh.sums;
    const int lvl = lm.lvl;
    const float warped = s[CH_GOOD] + s[CH_BAD];
    sh.action = ACT_CONTINUE;
    // MIN_GOODPERALL_PIXEL_ABSMIN is the float literal 0.01f (util/settings.h:170): the reference's threshold is a float product
    if (warped < 0.01f * (p.W >> lvl) * (p.H >> lvl)) {                          // :324-329 / :369-374
        lm.diverged = 1;
        sh.action = ACT_DIVERGED;
        return;
    }
    const float error = s[CH_SUMRESW] / warped;                                  // calcWeightsAndResidual, :789
    lm.nRes[lvl]++;
    const bool init = lm.phase == PH_INIT;
    const bool accept = init || error < lm.lastErr;                              // :381
    bool leave;
    if (accept) {
        if (!init) lm.refToFrame = lm.cand;
        if (p.useAffine) affineFromSums(s, lm.affine_a, lm.affine_b);            // :331-335 / :385-389
        const bool converged = !init && (error / lm.lastErr > p.st.convergenceEps[lvl]);   // :404
        if (!init) lm.last_residual = error;                                     // :414
        lm.lastErr = error;                                                      // :336 / :414
#pragma unroll
        for (int k = 0; k < 27; k++) lm.lsq[k] = s[k];                           // buffers now belong to this pose
        if (init) { lm.LM_lambda = p.st.lambdaInitial[lvl]; lm.iteration = 0; }  // :341
        else {
            if (lm.LM_lambda <= 0.2) lm.LM_lambda = 0;                           // :417-420
            else lm.LM_lambda *= p.st.lambdaSuccessFac;
            if (!converged) lm.iteration++;
        }
        leave = converged || !(lm.iteration < p.st.maxItsPerLvl[lvl]);           // :343, :411
        if (!leave) { lm.nUpd[lvl]++; lm.incTry = 0; }                           // calculateWarpUpdate(ls), :346
    } else {                                                                     // :424-447 reject
        float dot = 0;
#pragma unroll
        for (int i = 0; i < 6; i++) dot += lm.inc[i] * lm.inc[i];
        leave = !(dot > p.st.stepSizeMin[lvl]);                                  // :432-441
        if (!leave) {
            if (lm.LM_lambda == 0) lm.LM_lambda = 0.2;                           // :443-446
            else lm.LM_lambda *= pow((double)p.st.lambdaFailFac, lm.incTry);
        }
    }
    if (leave) lmNextLevel(p, lm, sh);
    else lmSolveAndPropose(lm, sh);
}

__global__ void __launch_bounds__(TP_THREADS, 1) k_track_persistent(const __grid_constant__ TrackParams p, TrackState* __restrict__ out, TrackState* __restrict__ outDev)
{
    __shared__ LMShared sh;
    __shared__ alignas(LMState) unsigned char lmStorage[sizeof(LMState)];     // every field is written before use; no constructor in shared memory
    LMState& lm = *reinterpret_cast<LMState*>(lmStorage);
    __shared__ float sm[TP_THREADS / 32][EV_NCH];
    static_assert(EV_NCH == 40, "warpReduceAcc is written for 32 + 8 channels");
    extern __shared__ __align__(128) unsigned char winSmem[];      // TP_WARPS windows of TRK_WIN_H x TRK_WIN_W float4
    __shared__ __align__(8) uint64_t winBar[TP_WARPS];
    __shared__ float xrow[2][EV_NCH];                              // this CTA's row for the cluster-local exchange
    unsigned int xphase = 0;
    const int clusterSize = (int)cg::this_cluster().num_blocks();
    unsigned int epoch = 0;
    long long cyc[6] = { 0, 0, 0, 0, 0, 0 };
    const long long tStart = clock64();
    WarpWindow W;
    W.win = reinterpret_cast<const float4*>(winSmem) + (size_t)(threadIdx.x >> 5) * (TRK_WIN_W * TRK_WIN_H);
    W.bar = &winBar[threadIdx.x >> 5];
    W.parity = 0u; W.lvl = -1; W.ox = 0; W.oy = 0; W.valid = false; W.hits = 0u; W.pcLvl = -1; W.pcMode = -1; W.pcIsPoint = false;
    W.pcx = W.pcy = W.pcz = W.pcVar = W.pcColor = 0.f;
    if (p.useTma && (threadIdx.x & 31) == 0) {
        mbarInit(W.bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }

    // fresh refPixelWasGood mask: all true (Frame.h:433); ordered before the level-1 evaluations by the barriers
    if (p.maskFresh) {
        uint32_t* m32 = reinterpret_cast<uint32_t*>(p.goodMask);
        for (int i = blockIdx.x * TP_THREADS + threadIdx.x; i < p.maskBytes / 4; i += gridDim.x * TP_THREADS) m32[i] = 0x01010101u;
    }

    if (threadIdx.x == 0) {
        for (int i = 0; i < 4; i++) lm.refToFrame.q[i] = p.initRefToFrame[i];
        for (int i = 0; i < 3; i++) lm.refToFrame.t[i] = p.initRefToFrame[4 + i];
        for (int l = 0; l < LSD_LEVELS; l++) { lm.nRes[l] = 0; lm.nUpd[l] = 0; }
        lm.affine_a = 1.f; lm.affine_b = 0.f; lm.lastErr = 0.f; lm.last_residual = 0.f; lm.LM_lambda = 0.f;
        lm.diverged = 0; lm.incTry = 0; lm.iteration = 0; lm.dbg[0] = lm.dbg[1] = lm.dbg[2] = lm.dbg[3] = 0;
        lm.lvl = SE3TRACKING_MAX_LEVEL - 1; lm.phase = PH_INIT;
        sh.lvl = lm.lvl; sh.action = ACT_CONTINUE;
        setEvalPose(sh.pose, lm.refToFrame, lm.affine_a, lm.affine_b);
    }
    __syncthreads();

    while (true) {
        const int lvl = sh.lvl;
        const int npx = p.lvl[lvl].w * p.lvl[lvl].h;
        const int evalMode = npx <= TP_LOCAL_MAX_PIXELS ? EVAL_CTA
                           : (clusterSize > 1 && npx <= p.clusterLocalMaxPixels) ? EVAL_CLUSTER : EVAL_GRID;
        gridEvaluate(p, lvl, evalMode, sh, sm, xrow, xphase, epoch, cyc, W);
        if (threadIdx.x == 0) { const long long t0 = clock64(); lmAdvance(p, lm, sh); lm.dbg[0] += clock64() - t0; }
        __syncthreads();
        if (sh.action != ACT_CONTINUE) break;
    }

    if (clusterSize > 1) cg::this_cluster().sync();             // nobody may exit while a peer can still read its xrow
    if (p.debug) {
        atomicAdd(p.barrier + 41, W.hits & 0xffffu);
        atomicAdd(p.barrier + 42, W.hits >> 16);
    }
    if (p.debug && threadIdx.x == 0 && blockIdx.x < TP_MAXGRID_DBG) {
        long long tot = clock64() - tStart;
        for (int i = 0; i < 4; i++) out->cycBlk[blockIdx.x][i] = cyc[i];
        out->cycBlk[blockIdx.x][5] = tot;
        out->cycBlk[blockIdx.x][4] = tot - cyc[0] - cyc[1] - cyc[2] - cyc[3];
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        lsdgpu_eval_result ev;
        evalFinish(sh.sums, &ev);        // statistics of the LAST EVALUATED pose
        for (int i = 0; i < 4; i++) out->refToFrame[i] = lm.refToFrame.q[i];
        for (int i = 0; i < 3; i++) out->refToFrame[4 + i] = lm.refToFrame.t[i];
        out->pointUsage = ev.pointUsage; out->goodCount = ev.goodCount; out->badCount = ev.badCount;
        out->meanRes = ev.meanRes; out->lastResidual = lm.last_residual;
        out->affine_a = lm.affine_a; out->affine_b = lm.affine_b;
        out->diverged = lm.diverged;
        for (int l = 0; l < LSD_LEVELS; l++) { out->numCalcResidualCalls[l] = lm.nRes[l]; out->numCalcWarpUpdateCalls[l] = lm.nUpd[l]; }
        out->totalEvals = (int)epoch;
        // device-resident copy of what the mapping kernels of the same frame need (no host round trip in between)
        for (int i = 0; i < 7; i++) outDev->refToFrame[i] = out->refToFrame[i];
        outDev->pointUsage = ev.pointUsage; outDev->goodCount = ev.goodCount; outDev->badCount = ev.badCount;
        outDev->lastResidual = lm.last_residual; outDev->diverged = lm.diverged;
        if (p.doPrepare)
            devicePrepareObserve(lm.refToFrame, lm.diverged, lm.last_residual, ev.pointUsage, ev.goodCount, ev.badCount, p.prep, p.obsOut, p.skipOut);
        cyc[5] = clock64() - tStart;
        cyc[4] = cyc[5] - cyc[0] - cyc[1] - cyc[2] - cyc[3];
        for (int i = 0; i < 6; i++) out->cyc[i] = cyc[i];
        if (p.debug) for (int i = 0; i < 3; i++) out->cycBlk[TP_MAXGRID_DBG - 1][i] = lm.dbg[i];
        __threadfence_system();
        out->doneSeq = p.launchSeq;
    }
}

static cudaError_t trackPersistentSetup(lsdgpu_ctx* ctx)
{
    int coop = 0;
    cudaError_t e = cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, ctx->device);
    if (e != cudaSuccess) return e;
    if (!coop) return cudaErrorNotSupported;
    e = cudaFuncSetAttribute((const void*)k_track_persistent, cudaFuncAttributeMaxDynamicSharedMemorySize, TP_WIN_SMEM);
    if (e != cudaSuccess) return e;
    // Optional (LSDGPU_TRACK_CLUSTER=2|4|8): launch as thread-block clusters and evaluate levels of <= 8192 pixels (L3, L4
    // at 640x480) redundantly per cluster with a DSMEM exchange instead of the grid barrier.  Measured on B200: SLOWER
    // (cluster 4: 228k vs 211k cycles per frame; cluster 2: 275k) -- a warp then walks 2-3 chunks sequentially and the
    // per-chunk dependency chain (~2 700 cycles) dominates; the grid has enough warps to give every chunk its own.
    // Default: no clusters.  The grid is the largest whole number of co-resident clusters (B200: 33 x 4 = 132 CTAs).
    {
        const char* ce = getenv("LSDGPU_TRACK_CLUSTER");
        int cs = ce ? atoi(ce) : 1;
        if (cs != 1 && cs != 2 && cs != 4 && cs != 8) cs = 1;
        ctx->trackCluster = 1;
        ctx->trackGrid = ctx->smCount < TP_MAXGRID ? ctx->smCount : TP_MAXGRID;
        if (cs > 1) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(ctx->smCount / cs * cs); cfg.blockDim = dim3(TP_THREADS); cfg.dynamicSmemBytes = TP_WIN_SMEM;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            int nc = 0;
            if (cudaOccupancyMaxActiveClusters(&nc, (const void*)k_track_persistent, &cfg) == cudaSuccess && nc * cs >= 64) {
                ctx->trackCluster = cs;
                ctx->trackGrid = nc * cs < TP_MAXGRID ? nc * cs : (TP_MAXGRID / cs) * cs;
            } else
                cudaGetLastError();
        }
    }
    if (getenv("LSDGPU_TRACK_DEBUG")) {
        fprintf(stderr, "[track] cluster %d, grid %d\n", ctx->trackCluster, ctx->trackGrid);
        for (int cs = 1; cs <= 8; cs *= 2) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(ctx->smCount / cs * cs); cfg.blockDim = dim3(TP_THREADS); cfg.dynamicSmemBytes = TP_WIN_SMEM;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            int nc = -1;
            cudaError_t ce = cudaOccupancyMaxActiveClusters(&nc, (const void*)k_track_persistent, &cfg);
            fprintf(stderr, "[track] cluster size %d: max active clusters %d (%s) -> %d CTAs\n", cs, nc, cudaGetErrorString(ce), nc * cs);
        }
        cudaGetLastError();
    }
    return cudaSuccess;
}

static int trackPersistentFinish(lsdgpu_ctx* ctx, FrameSlot* fr, lsdgpu_track_result* out);

// device time of the last profiled tracking launch (CUDA events on the context's stream), accumulated lazily
static void flushTrackProfile(lsdgpu_ctx* ctx)
{
    if (!ctx->trackProfilePending) return;
    float ms = 0;
    cudaEventSynchronize(ctx->kEnd);
    if (cudaEventElapsedTime(&ms, ctx->kBegin, ctx->kEnd) == cudaSuccess) { ctx->trackKernelMs += ms; ctx->trackKernelLaunches++; }
    ctx->trackProfilePending = false;
}

// enqueue the tracking kernel of one frame on the context's stream (no host synchronisation)
static int trackPersistentEnqueue(lsdgpu_ctx* ctx, FrameSlot* kf, FrameSlot* fr, const double init_qt[7],
                                  const lsdgpu_track_settings* st, const PrepareConsts* prep = nullptr)
{
    flushTrackProfile(ctx);
    TrackParams P;
    memset(&P, 0, sizeof(P));
    for (int l = SE3TRACKING_MIN_LEVEL; l < SE3TRACKING_MAX_LEVEL; l++) {
        const LevelCam& c = ctx->cam[l];
        TrackLevelParams& L = P.lvl[l];
        L.kfIdepth = kf->idepth[l]; L.kfVar = kf->idepthVar[l]; L.kfColor = kf->image[l]; L.frameGrad = fr->grad[l];
        L.w = c.w; L.h = c.h; L.fx = c.fx; L.fy = c.fy; L.cx = c.cx; L.cy = c.cy;
        L.fxi = c.fxi; L.fyi = c.fyi; L.cxi = c.cxi; L.cyi = c.cyi;
    }
    for (int l = 0; l < LSD_LEVELS; l++) P.gradMap[l] = fr->gradMap[l];
    { const char* tm = getenv("LSDGPU_TRACK_TMA"); P.useTma = tm ? atoi(tm) : 1; }
    P.minLevel = SE3TRACKING_MIN_LEVEL;
    P.goodMask = fr->goodMask;
    P.maskFresh = fr->hasGoodMask ? 0 : 1;
    P.maskBytes = (ctx->w * ctx->h) / 4;
    P.W = ctx->w; P.H = ctx->h;
    lsd::SE3<double> init;
    for (int i = 0; i < 4; i++) init.q[i] = init_qt[i];
    for (int i = 0; i < 3; i++) init.t[i] = init_qt[4 + i];
    lsd::SE3<float> r2f = lsd::se3Cast<float>(lsd::se3Inverse(init));             // SE3Tracker.cpp:306
    for (int i = 0; i < 4; i++) P.initRefToFrame[i] = r2f.q[i];
    for (int i = 0; i < 3; i++) P.initRefToFrame[4 + i] = r2f.t[i];
    P.st = *st;
    P.C.cameraPixelNoise2 = ctx->g.cameraPixelNoise2; P.C.var_weight = st->var_weight; P.C.huber_half = st->huber_d / 2;
    P.useAffine = ctx->g.useAffineLightningEstimation;
    P.partials = ctx->evPartials;
    P.barrier = ctx->evCounter;
    { const char* bm = getenv("LSDGPU_BARRIER_MODE"); P.barrierMode = bm ? atoi(bm) : 1; }
    if (prep) { P.doPrepare = 1; P.prep = *prep; P.obsOut = ctx->dObs; P.skipOut = ctx->dSkipFlag; }
    TrackState* dOut = (TrackState*)ctx->dTrackStateMapped;
    TrackState* dOutDev = (TrackState*)ctx->dTrackState;

    const int grid = ctx->trackGrid;                                               // one CTA per SM (whole clusters)
    P.clusterLocalMaxPixels = ctx->trackCluster > 1 ? TP_CLUSTER_MAX_PIXELS : 0;
    void* args[] = { (void*)&P, (void*)&dOut, (void*)&dOutDev };
    const bool dbg = getenv("LSDGPU_TRACK_DEBUG") != nullptr;
    P.debug = dbg ? 1 : 0;
    P.barrierBase = ctx->barrierBase;
    P.launchSeq = ++ctx->trackSeq;
    if (dbg) LSD_CHECK(ctx, cudaMemsetAsync(ctx->evCounter + 40, 0, 4 * sizeof(unsigned int), ctx->stream));
    if (ctx->profileTrackKernel) cudaEventRecord(ctx->kBegin, ctx->stream);
    {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid); cfg.blockDim = dim3(TP_THREADS); cfg.dynamicSmemBytes = TP_WIN_SMEM; cfg.stream = ctx->stream;
        cudaLaunchAttribute at[2];
        at[0].id = cudaLaunchAttributeCooperative; at[0].val.cooperative = 1;
        at[1].id = cudaLaunchAttributeClusterDimension;
        at[1].val.clusterDim.x = ctx->trackCluster; at[1].val.clusterDim.y = 1; at[1].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = ctx->trackCluster > 1 ? 2 : 1;
        LSD_CHECK(ctx, cudaLaunchKernelExC(&cfg, (const void*)k_track_persistent, args));
    }
    ctx->launches++;
    if (ctx->profileTrackKernel) cudaEventRecord(ctx->kEnd, ctx->stream);
    fr->hasGoodMask = true;
    ctx->trackUseTma = P.useTma;
    return 0;
}

static int trackPersistent(lsdgpu_ctx* ctx, FrameSlot* kf, FrameSlot* fr, const double init_qt[7],
                           const lsdgpu_track_settings* st, lsdgpu_track_result* out)
{
    int r = trackPersistentEnqueue(ctx, kf, fr, init_qt, st);
    if (r) return r;
    return trackPersistentFinish(ctx, fr, out);
}

// wait for the frame's work and translate the mapped result block (SE3Tracker.cpp:453-485)
static int trackPersistentFinish(lsdgpu_ctx* ctx, FrameSlot* fr, lsdgpu_track_result* out)
{
    memset(out, 0, sizeof(*out));
    TrackState* hOut = (TrackState*)ctx->hTrackState;
    const int grid = ctx->trackGrid;
    // The result block lives in mapped pinned memory and its last word is a sequence number: spin on it (a few
    // hundred ns of latency) instead of a stream synchronisation.  Later launches are stream-ordered anyway.
    if (getenv("LSDGPU_TRACK_DEBUG")) LSD_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    else {
        bool done = false;
        for (long long spin = 0; spin < 200000000LL; spin++) {
            if (hOut->doneSeq == ctx->trackSeq) { done = true; break; }
            if ((spin & 0xffff) == 0xffff && cudaStreamQuery(ctx->stream) != cudaErrorNotReady) { done = hOut->doneSeq == ctx->trackSeq; break; }
        }
        if (!done) LSD_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    }
    // the other fields of the mapped result block are read with plain loads below: keep them behind the doneSeq read.  Block 0
    // publishes doneSeq while other CTAs may still be exiting, so only STREAM-ORDERED consumers may assume the kernel has retired.
    std::atomic_thread_fence(std::memory_order_acquire);
    ctx->barrierBase += (unsigned int)hOut->totalEvals * (unsigned int)grid;

    if (ctx->profileTrackKernel) {           // the event pair is read lazily (flushTrackProfile): no extra sync on the path
        double bytes = 0;
        for (int l = SE3TRACKING_MIN_LEVEL; l < SE3TRACKING_MAX_LEVEL; l++)
            bytes += (double)hOut->numCalcResidualCalls[l] * ((double)ctx->cam[l].w * ctx->cam[l].h * (12.0 + 16.0 + (l == 1 ? 1.0 : 0.0)) + EV_NCH * 4.0);
        ctx->trackKernelBytes += bytes;
        ctx->trackProfilePending = true;
    }

    if (getenv("LSDGPU_TRACK_DEBUG")) {
        unsigned int dbgc[3] = { 0, 0, 0 };
        cudaMemcpy(dbgc, ctx->evCounter + 40, 12, cudaMemcpyDeviceToHost);
        fprintf(stderr, "[track] useTma=%d tmaTimeouts=%u taps: %u from the smem window, %u through L1/L2\n", ctx->trackUseTma, dbgc[0], dbgc[1], dbgc[2]);
        fprintf(stderr, "[track] evals=%d cycles: points=%lld ctaReduce=%lld barrier=%lld combine=%lld serialLM=%lld total=%lld\n",
                hOut->totalEvals, hOut->cyc[0], hOut->cyc[1], hOut->cyc[2], hOut->cyc[3], hOut->cyc[4], hOut->cyc[5]);
        fprintf(stderr, "   thread0: lmAdvance=%lld (solve=%lld pose=%lld)\n", hOut->cycBlk[TP_MAXGRID_DBG - 1][0], hOut->cycBlk[TP_MAXGRID_DBG - 1][1], hOut->cycBlk[TP_MAXGRID_DBG - 1][2]);
        const char* nm[6] = { "points", "ctaReduce", "barrier", "combine", "serial", "total" };
        for (int k = 0; k < 6; k++) {
            long long mn = 1LL << 60, mx = 0, sum = 0; int imx = 0, imn = 0;
            for (int b = 0; b < grid && b < TP_MAXGRID_DBG; b++) {
                long long v = hOut->cycBlk[b][k];
                if (v < mn) { mn = v; imn = b; }
                if (v > mx) { mx = v; imx = b; }
                sum += v;
            }
            fprintf(stderr, "   %-9s min=%lld (cta %d) max=%lld (cta %d) mean=%lld\n", nm[k], mn, imn, mx, imx, sum / grid);
        }
    }
    out->pointUsage = hOut->pointUsage; out->lastGoodCount = hOut->goodCount; out->lastBadCount = hOut->badCount;
    out->lastMeanRes = hOut->meanRes;
    out->affineEstimation_a = hOut->affine_a; out->affineEstimation_b = hOut->affine_b;
    for (int l = 0; l < LSD_LEVELS; l++) {
        out->numCalcResidualCalls[l] = hOut->numCalcResidualCalls[l];
        out->numCalcWarpUpdateCalls[l] = hOut->numCalcWarpUpdateCalls[l];
    }
    if (hOut->diverged) {
        out->frameToRef_qt[3] = 1;
        out->diverged = 1; out->trackingWasGood = 0;
        return 0;
    }
    const int W = ctx->w, H = ctx->h;
    out->lastResidual = hOut->lastResidual;
    out->trackingWasGood = hOut->goodCount / ((W >> SE3TRACKING_MIN_LEVEL) * (H >> SE3TRACKING_MIN_LEVEL)) > 0.04f
                           && hOut->goodCount / (hOut->goodCount + hOut->badCount) > 0.5f;
    out->initialTrackedResidual = out->lastResidual / out->pointUsage;
    lsd::SE3<float> T;
    for (int i = 0; i < 4; i++) T.q[i] = hOut->refToFrame[i];
    for (int i = 0; i < 3; i++) T.t[i] = hOut->refToFrame[4 + i];
    lsd::SE3<double> f2r = lsd::se3Cast<double>(lsd::se3Inverse(T));
    for (int i = 0; i < 4; i++) out->frameToRef_qt[i] = f2r.q[i];
    for (int i = 0; i < 3; i++) out->frameToRef_qt[4 + i] = f2r.t[i];
    return 0;
}
