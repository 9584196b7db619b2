// dec_cluster.cu — one launch for ALL decoder layers of a decode step (bf16 build, whisper-base widths):
// replaces the 48 per-layer launches (qkv / self-attention / o / cq / cross-attention / co / fc1 / fc2 x 6)
// that made the round-1 decode step latency-bound (VERDICT r1 "weak" 3).  Reference: the body of one
// decoder_with_past_model.onnx run, /root/reference/src/main.rs:793-826 (binding.run at :814).
//
// Shape of the kernel.  Sequences are independent, so the batch is dealt over thread-block clusters of 16 CTAs
// (8 where the part cannot schedule 16): a B200 holds 7 such clusters at once (its GPCs have 16+ SMs, one has
// fewer), so a batch of 32 becomes 7 clusters of 5/5/5/5/4/4/4 sequences on 112 SMs.  A cluster walks the 8
// stages of every layer on its own; the only synchronisation between CTAs is the hardware cluster barrier
// (release/acquire) between stages — no grid-wide barrier, no co-residency requirement between clusters, so any
// number of these kernels (batches in flight) can share the GPU, and clusters drift against each other: one
// cluster's weight phase overlaps another's K/V streaming.
//   * every GEMM stage splits the weight ROWS over the CTAs of the cluster, the cluster's sequences form one n-tile
//     of mma.sync m16n8k16 (weights are the M side: swap-AB, no batch padding waste), a warp owns one m-tile of 16
//     rows with the whole contraction in four interleaved accumulators; activations are exchanged through L2
//     (ld.cg after the barrier), each CTA keeps its own columns of the residual stream and its bias slices in
//     shared memory;
//   * the attentions run on the tensor cores as well: S = K q and O = V^T p as m16n8k16 with the K / V tile as the A
//     operand (ldmatrix, .trans for V) and q / p as a two-column B operand (bf16 high part + bf16 residual, so q
//     and p keep ~16 mantissa bits).  The cross-attention (the dominant HBM stream: 2 x 1500 x 64 bf16 per
//     (sequence, head) and layer) is cut into half-units (one head of one sequence, keys 0..767 or 768..1499) dealt
//     evenly over the CTAs; a half-unit is TWO passes — all its K tiles (scores to shared memory, every tile an
//     independent mma chain), one softmax over the 768 scores, all its V tiles — instead of an online softmax whose
//     max / rescale / exp chain serialised every tile (measured: ~1000 cycles per 128-key tile per SM).  The consumer
//     stage merges the two halves.  The cut never depends on where a sequence sits in the batch, so a clip's tokens
//     do not depend on its batch position either.
// Everything a CTA reads from HBM — its weight slices AND its K/V tiles — flows through ONE shared-memory ring of
// DC_NSLOT x 17 KB slots (a weight m-tile of 16 rows x 512 k out of a pitched copy of the weights as one
// cp.async.bulk; a K or a V tile of 128 keys x 64 dims as one 128B-swizzled tensor-map tile) with a full and an
// empty mbarrier per slot.  The order of tiles is a static schedule; lane 0 of a ninth, producer-only warp walks
// it, refilling a slot as soon as its readers released it, so the next stages' weights and the next K/V tiles are
// in flight while the compute warps sit in a barrier or a latency-bound stage.  The producer warp follows the same
// control flow as the compute warps (it takes part in every cluster barrier) but never computes and never joins
// the compute warps' named barriers.
#include <cuda.h>

#include "ctx.h"

namespace {

using bf16 = __nv_bfloat16;

constexpr int DC_D = 512, DC_H = 8, DC_FFN = 2048, DC_HD = 64;    // whisper-base decoder widths
constexpr int DC_CT = 256, DC_THREADS = DC_CT + 32;               // 8 compute warps + 1 producer warp (tile issue only)
constexpr int DC_CROWS = 16, DC_KC = 512;                          // weight chunk: one m-tile, 16 rows x 512 k (bf16)
constexpr int DC_WPITCH = DC_KC + 32;                              // elements per stored row: 1088 B, conflict-free 128-bit reads
constexpr int DC_WROW = DC_WPITCH * 2;
constexpr int DC_SLOT = DC_CROWS * DC_WROW;                        // 17408 B = one weight chunk image (a K or V tile uses 16384)
constexpr int DC_NSLOT = 10;
constexpr int DC_KEYS = 128;                                       // keys per K / V tile
constexpr int DC_TILE = DC_KEYS * DC_HD * 2;                       // 16 KB
constexpr int DC_SEQ = 8;                                          // sequences per cluster <= one mma n-tile
constexpr int DC_XS = DC_SEQ * (DC_FFN * 2 + 64);                  // staged activations, widest stage (fc2)
constexpr int DC_MAXHALF = 6;                                      // K (or V) tiles of one attention pass: 768 keys
constexpr int DC_MAXTILES = 192;                                   // schedule entries of one layer (per CTA)
constexpr int OFF_XS = DC_NSLOT * DC_SLOT;
constexpr int OFF_PART = OFF_XS + DC_XS;                           // 8 KB: fc2 partials | attention: scores, p_hi, p_lo, staged q|k|v
constexpr int OFF_ACC = OFF_PART + 8192;                           // attention merge scratch [8][64] + red[64]
constexpr int OFF_TAB = OFF_ACC + (8 * 64 + 64) * 4;               // schedule table
constexpr int DC_NBIAS = (3 * DC_D + 3 * DC_D + DC_FFN + DC_D) / 16, DC_BIAS_LAYERS = 6;    // bias floats per layer of a CTA (cluster of 16)
constexpr int OFF_BIAS = OFF_TAB + DC_MAXTILES * 8;                // [layers <= 6][DC_NBIAS] this CTA's bias slices
constexpr int OFF_XOWN = OFF_BIAS + DC_BIAS_LAYERS * DC_NBIAS * 4; // [64 rows][8 seqs] this CTA's slice of the residual stream
constexpr int OFF_BAR = OFF_XOWN + 64 * 8 * 4;
constexpr int DC_SMEM = OFF_BAR + 2 * DC_NSLOT * 8 + 16;
static_assert(DC_SMEM <= 232448, "over the 227 KB of shared memory a CTA can opt into");
// weight image of one layer: every matrix as [N/16][K/512] chunks of [16 rows][544] bf16
constexpr int IMG_QKV = 0, IMG_O = IMG_QKV + 3 * DC_D / DC_CROWS, IMG_CQ = IMG_O + DC_D / DC_CROWS, IMG_CO = IMG_CQ + DC_D / DC_CROWS,
              IMG_FC1 = IMG_CO + DC_D / DC_CROWS, IMG_FC2 = IMG_FC1 + DC_FFN / DC_CROWS,
              IMG_CHUNKS = IMG_FC2 + (DC_D / DC_CROWS) * (DC_FFN / DC_KC);
constexpr float DC_QSCALE = 0.125f * 1.4426950408889634f;         // head_dim^-1/2 * log2(e): scores live in the log2 domain (ex2)

struct DcLayer {                    // device-side pointer table of one decoder layer (small f32 vectors only)
    const float *bqkv, *bo, *bcq, *bco, *bfc1, *bfc2;
    const float *ln1w, *ln1b, *ln2w, *ln2b, *ln3w, *ln3b;
};
struct DcArgs {
    const int* state;               // [0] = position s of the token fed this step, [1] = prompt_len
    const int* prompt;
    const int* cur_tok;
    const bf16* E;                  // [vocab][d]
    const float* P;                 // [n_text_ctx][d]
    const DcLayer* layers;
    const unsigned char* wimg;      // [layers][IMG_CHUNKS][DC_SLOT] pitched weight chunks
    int n_layers;
    float* x;                       // [B][d] residual stream (output: input of the final LayerNorm)
    float* qkv;                     // [B][3d]
    float* q;                       // [B][d]
    bf16* att;                      // [B][d] self-attention output
    float* xpart;                   // [B][H][2][66] cross-attention states of the two key halves (acc[64], m (log2 domain), l)
    bf16* ffn;                      // [B][ffn]
    bf16* self_kv;                  // [layers][Bmax][T_max][2d]
    int B, Bmax, T_max, Tk, n_clusters;
    int prof_stage;                 // which stage of layer 1 gets the fine stamps (5 = cross-attention, 7 = fc1, ...)
    long long* prof;                // optional: clock64 stamps of cluster 0 / CTA 0 at every stage boundary (WB_DEC_PROF=1)
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive_n(uint32_t bar, uint32_t n) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(n) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try(uint32_t bar, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n.reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, 100000;\n"
        "selp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    return done;
}
// Bounded wait: a schedule bug must end in a trap (launch error), never in a hung GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    int spins = 0;
    while (!mbar_try(bar, parity))
        if (++spins > 20000) __trap();
}
__device__ __forceinline__ void bulk_copy(void* dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma_tile_3d(const CUtensorMap* map, uint32_t bar, void* dst, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
// 16-byte LDGSTS; src_bytes = 0 writes zeros (rows past the end of the vocabulary)
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
// this thread's earlier cp.async count as ONE of the barrier's expected arrivals once they have all landed
__device__ __forceinline__ void cp_async_arrive(uint32_t bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ int cluster_ctarank() { int r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ int cluster_id_x() { int r; asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r)); return r; }
__device__ __forceinline__ void cbar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }      // the 8 compute warps only

__device__ __forceinline__ void mma_bf16(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float bf16_round(float x) { return __bfloat162float(__float2bfloat16(x)); }
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }   // 2^x, 2^-inf = 0

// ---- the tile stream of one CTA ----
enum { T_W = 0, T_CROSS = 1, T_SELF = 2 };
// T_W: a = chunk index in the layer image;  K/V tiles: a = kind<<24 | v<<23 | h<<16 | key0, b = sequence
struct TileEnt { uint32_t a, b; };

template <int CS>
struct Stream {
    static constexpr int RQ = 3 * DC_D / CS, RO = DC_D / CS, RF = DC_FFN / CS;       // weight rows of this CTA per matrix
    static constexpr int NQ = RQ / DC_CROWS, NO = RO / DC_CROWS, NF1 = RF / DC_CROWS, KQ2 = DC_FFN / DC_KC, NF2 = NO * KQ2;
    static_assert(RO % DC_CROWS == 0 && RQ % DC_CROWS == 0 && RF % DC_CROWS == 0, "row slices are whole chunks");
    static_assert((2 * DC_H) % CS == 0 || CS % (2 * DC_H) == 0, "half-units deal evenly");
    const DcArgs& a;
    const CUtensorMap* tm_cross;
    const CUtensorMap* tm_self;
    unsigned char* smem;
    TileEnt* tab;
    uint32_t full, empty;           // mbarrier arrays (shared-memory addresses)
    int rank, b0, nb;               // this cluster: sequences [b0, b0 + nb)
    int n_self_units, nst = 0;      // self-attention: units rank, rank+CS, ... ; K (= V) tiles of cached keys per unit
    int tph, nhu, ncross;           // cross-attention: K (= V) tiles per half-unit, half-units of this CTA, tiles of this CTA
    int TL = 0, total = 0;
    int cons = 0, issued = 0, il = 0, ii = 0;   // consumed (program order) / issued tiles; (layer, entry) of the next tile to issue
    int warp, lane;
    int n2 = 1 << 30;               // fine stamps (WB_DEC_PROF=1): armed for the cross-attention of layer 1
    __device__ __forceinline__ void stamp2() {
        if (n2 < 256 + 200) {
            if (threadIdx.x == 0 && blockIdx.x == 0 && a.prof != nullptr) a.prof[n2] = clock64();
            ++n2;
        }
    }

    __device__ __forceinline__ Stream(const DcArgs& a_, const CUtensorMap* tc, const CUtensorMap* ts, unsigned char* sm)
        : a(a_), tm_cross(tc), tm_self(ts), smem(sm) {
        warp = threadIdx.x >> 5; lane = threadIdx.x & 31;
        tab = reinterpret_cast<TileEnt*>(sm + OFF_TAB);
        full = smem_u32(sm + OFF_BAR);
        empty = full + DC_NSLOT * 8;
        rank = cluster_ctarank();
        const int c = cluster_id_x(), base = a.B / a.n_clusters, rem = a.B % a.n_clusters;
        nb = base + (c < rem ? 1 : 0);
        b0 = c * base + min(c, rem);
        tph = ((a.Tk + DC_KEYS - 1) / DC_KEYS) / 2;                // host guarantees an even tile count
        const int units = nb * DC_H;
        n_self_units = rank < units ? (units - rank + CS - 1) / CS : 0;
        nhu = units * 2 / CS;                                      // 16 half-units per sequence over CS CTAs
        ncross = nhu * 2 * tph;
    }
    __device__ __forceinline__ static uint32_t chunk_of(int img0, int rows_per_cta, int KQ, int rank, int rc, int kq) {
        return (uint32_t)(img0 + (rank * (rows_per_cta / DC_CROWS) + rc) * KQ + kq);
    }
    __device__ __forceinline__ static TileEnt kv_ent(int kind, int v, int h, int key0, int b) {
        return TileEnt{((uint32_t)kind << 24) | ((uint32_t)v << 23) | ((uint32_t)h << 16) | (uint32_t)key0, (uint32_t)b};
    }
    __device__ __forceinline__ TileEnt entry(int i) const {
        if (i < NQ) return TileEnt{chunk_of(IMG_QKV, RQ, 1, rank, i, 0), 0u};
        i -= NQ;
        const int nself = n_self_units * 2 * nst;                  // per unit: its K tiles, then its V tiles
        if (i < nself) {
            const int u = i / (2 * nst), j = i - u * 2 * nst, unit = rank + u * CS;
            return kv_ent(T_SELF, j >= nst, unit % DC_H, (j % nst) * DC_KEYS, b0 + unit / DC_H);
        }
        i -= nself;
        if (i < NO) return TileEnt{chunk_of(IMG_O, RO, 1, rank, i, 0), 0u};
        i -= NO;
        if (i < NO) return TileEnt{chunk_of(IMG_CQ, RO, 1, rank, i, 0), 0u};
        i -= NO;
        if (i < ncross) {                                           // per half-unit: its K tiles, then its V tiles
            const int hi = i / (2 * tph), j = i - hi * 2 * tph, hu = rank * nhu + hi, unit = hu >> 1, half = hu & 1;
            return kv_ent(T_CROSS, j >= tph, unit % DC_H, (half * tph + j % tph) * DC_KEYS, b0 + unit / DC_H);
        }
        i -= ncross;
        if (i < NO) return TileEnt{chunk_of(IMG_CO, RO, 1, rank, i, 0), 0u};
        i -= NO;
        if (i < NF1) return TileEnt{chunk_of(IMG_FC1, RF, 1, rank, i, 0), 0u};
        i -= NF1;
        return TileEnt{chunk_of(IMG_FC2, RO, KQ2, rank, i / KQ2, i % KQ2), 0u};
    }
    // the first NQ tiles (q|k|v weights) do not depend on the step: they are issued before the PDL wait, the rest of
    // the schedule needs the position s (number of cached self-attention keys)
    __device__ __forceinline__ void build_table(int s) {
        nst = (s + DC_KEYS - 1) / DC_KEYS;                          // cached keys 0..s-1 (the new key comes from registers)
        TL = NQ + n_self_units * 2 * nst + 3 * NO + ncross + NF1 + NF2;
        if (TL > DC_MAXTILES || nst > DC_MAXHALF || tph > DC_MAXHALF) __trap();
        total = TL * a.n_layers;
        for (int i = threadIdx.x; i < TL; i += DC_THREADS) tab[i] = entry(i);
    }
    // producer lane: start the loads of the next tile of the schedule
    __device__ __forceinline__ void issue_next() {
        const int slot = issued % DC_NSLOT;
        if (issued >= DC_NSLOT) mbar_wait(empty + slot * 8, (uint32_t)(issued / DC_NSLOT - 1) & 1u);      // its last readers are done
        unsigned char* dst = smem + slot * DC_SLOT;
        const uint32_t bar = full + slot * 8;
        const TileEnt e = (TL == 0) ? TileEnt{chunk_of(IMG_QKV, RQ, 1, rank, ii, 0), 0u} : tab[ii];
        const uint32_t kind = e.a >> 24;
        if (kind == T_W) {
            mbar_expect_tx(bar, DC_SLOT);
            bulk_copy(dst, a.wimg + ((size_t)il * IMG_CHUNKS + e.a) * DC_SLOT, DC_SLOT, bar);
        } else {
            // one TMA tensor tile ([128 keys][64 dims], 128B swizzle, rows past the cache end zero-filled by the unit;
            // self-attention rows >= s hold older decodes' data and are masked by the consumer)
            const int v = (e.a >> 23) & 1, h = (e.a >> 16) & 0x7f, key0 = e.a & 0xffff;
            mbar_expect_tx(bar, DC_TILE);
            tma_tile_3d(kind == T_CROSS ? tm_cross : tm_self, bar, dst, v * DC_D + h * DC_HD, key0, il * a.Bmax + (int)e.b);
        }
    }
    // every thread calls this after advancing `cons` (program order); only the producer lane acts: it keeps the ring
    // DC_NSLOT tiles ahead of the program point, blocking on the empty barriers (never on a CTA barrier)
    __device__ __forceinline__ void produce() {
        if (warp != 8) return;                                     // compute warps only track `cons`
        const int limit = TL == 0 ? NQ : total;                    // before the table exists: the static prefix only
        const int want = min(cons + DC_NSLOT, limit);
        while (issued < want) {
            if (threadIdx.x == DC_CT) issue_next();                // lane 0 of the producer warp
            ++issued;
            if (++ii == TL) { ii = 0; ++il; }                      // TL == 0 (no table yet): never wraps
        }
    }
    __device__ __forceinline__ unsigned char* slot_of(int t) const { return smem + (t % DC_NSLOT) * DC_SLOT; }
    __device__ __forceinline__ void wait_full(int t) { mbar_wait(full + (t % DC_NSLOT) * 8, (uint32_t)(t / DC_NSLOT) & 1u); }
    // n <= 8 consecutive tiles: compute warp w waits for tile t + w (a probe of an mbarrier costs ~170 cycles even when the
    // phase is complete, and they do not overlap inside a warp); the caller's next cbar() makes all of them visible to all
    __device__ __forceinline__ void wait_full_spread(int t, int n) {
        if (warp < n) wait_full(t + warp);
    }
    // this warp is done reading tile t; n = arrivals it stands for (8 when it was the tile's only reader)
    __device__ __forceinline__ void release(int t, int n) {
        __syncwarp();
        if (lane == 0) mbar_arrive_n(empty + (t % DC_NSLOT) * 8, (uint32_t)n);
    }
};

// ---- activation staging: [8 sequences][K] bf16 rows (stride K*2+64) for the mma B operand (compute warps only) ----
// LayerNorm of an f32 row held across the lanes of one warp (two-pass statistics, like the oracle).
__device__ __forceinline__ void ln_pack_row(float4 (&v)[4], const float4 (&gw)[4], const float4 (&gb)[4], unsigned char* xrow, int lane) {
    float s1 = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) s1 += (v[i].x + v[i].y) + (v[i].z + v[i].w);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    const float mean = s1 / (float)DC_D;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float t0 = v[i].x - mean, t1 = v[i].y - mean, t2 = v[i].z - mean, t3 = v[i].w - mean;
        q += (t0 * t0 + t1 * t1) + (t2 * t2 + t3 * t3);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rs = 1.0f / sqrtf(q / (float)DC_D + 1e-5f);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int c = i * 128 + lane * 4;
        uint2 pk;
        pk.x = pack_bf16((v[i].x - mean) * rs * gw[i].x + gb[i].x, (v[i].y - mean) * rs * gw[i].y + gb[i].y);
        pk.y = pack_bf16((v[i].z - mean) * rs * gw[i].z + gb[i].z, (v[i].w - mean) * rs * gw[i].w + gb[i].w);
        *reinterpret_cast<uint2*>(xrow + c * 2) = pk;
    }
}
__device__ __forceinline__ void ln_params(const float* __restrict__ lw, const float* __restrict__ lb, float4 (&gw)[4], float4 (&gb)[4], int lane) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        gw[i] = *reinterpret_cast<const float4*>(lw + i * 128 + lane * 4);
        gb[i] = *reinterpret_cast<const float4*>(lb + i * 128 + lane * 4);
    }
}
__device__ __forceinline__ void zero_row(unsigned char* xrow, int K, int lane) {
    for (int c = lane * 8; c < K; c += 256) *reinterpret_cast<uint4*>(xrow + c * 2) = make_uint4(0u, 0u, 0u, 0u);
}
// warp w (< 8) stages sequence w: LN(X[b0+w]) (X f32 [B][d], written by the other CTAs of the cluster: L2 reads)
__device__ __forceinline__ void stage_ln(const float* X, int b0, int nb, const float4 (&gw)[4], const float4 (&gb)[4], unsigned char* xs) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char* xrow = xs + warp * (DC_D * 2 + 64);
    if (warp < nb) {
        float4 v[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] = __ldcg(reinterpret_cast<const float4*>(X + (size_t)(b0 + warp) * DC_D + i * 128 + lane * 4));
        ln_pack_row(v, gw, gb, xrow, lane);
    } else {
        zero_row(xrow, DC_D, lane);
    }
}
// bf16 rows copied as they are (attention output / GELU(fc1) written as bf16 by their producers)
__device__ __forceinline__ void stage_bf16(const bf16* X, int K, int b0, int nb, unsigned char* xs) {
    const int per_row = K / 8, xstride = K * 2 + 64;
    for (int idx = threadIdx.x; idx < DC_SEQ * per_row; idx += DC_CT) {
        const int row = idx / per_row, c8 = idx - row * per_row;
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (row < nb) v = __ldcg(reinterpret_cast<const uint4*>(X + (size_t)(b0 + row) * K + c8 * 8));
        *reinterpret_cast<uint4*>(xs + row * xstride + c8 * 16) = v;
    }
}

// one weight m-tile (16 rows x 512 k at A) against the staged activations (8 sequences x K, k offset koff): four
// interleaved accumulators, 8 dependent mma each instead of one chain of 32
__device__ __forceinline__ void mma_mtile(const unsigned char* A, const unsigned char* xs, int xstride, int koff, float (&out)[4], int lane) {
    const int g = lane >> 2, t4 = lane & 3;
    const unsigned char* a0 = A + g * DC_WROW + 8 * t4 * 2;
    const unsigned char* bx = xs + g * xstride + (koff + 8 * t4) * 2;
    float acc[4][4] = {};
#pragma unroll
    for (int c = 0; c < DC_KC / 32; ++c) {
        const uint4 wa = *reinterpret_cast<const uint4*>(a0 + c * 64);
        const uint4 wb = *reinterpret_cast<const uint4*>(a0 + 8 * DC_WROW + c * 64);
        const uint4 xb = *reinterpret_cast<const uint4*>(bx + c * 64);
        mma_bf16(acc[(2 * c) & 3], wa.x, wb.x, wa.y, wb.y, xb.x, xb.y);
        mma_bf16(acc[(2 * c + 1) & 3], wa.z, wb.z, wa.w, wb.w, xb.z, xb.w);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) out[i] = (acc[0][i] + acc[1][i]) + (acc[2][i] + acc[3][i]);
}

// ---- GEMM stages, K = 512 (q|k|v, o, cq, co, fc1): warp w takes m-tile w of every pass of 8; no partial sums, no CTA
// barrier.  `pre(row,seq)` loads what the epilogue adds (bias, residual) before the tile is waited for, `fin(row,seq,v)`
// finishes a value (activation, own-slice bookkeeping), `put(row4,seq,v4)` stores four consecutive rows of one sequence:
// the 16 x 8 results of a warp are transposed through shared memory so that a sequence's outputs leave as 16-byte
// stores (the cluster barrier's release waits for every outstanding store: 1024 scattered 2-4 byte stores per CTA cost
// ~2000 cycles there).  row = row inside this CTA's slice of the matrix.
template <int CS, typename Pre, typename Fin, typename Put>
__device__ __forceinline__ void gemm_rows(Stream<CS>& S, int n_mt, const unsigned char* xs, Pre pre, Fin fin, Put put) {
    const int warp = S.warp, lane = S.lane, g = lane >> 2, t4 = lane & 3;
    float* tr = reinterpret_cast<float*>(S.smem + OFF_PART) + warp * (8 * 17);      // this warp's [8 seqs][16 rows (+1)] transpose tile
    for (int c0 = 0; c0 < n_mt; c0 += 8) {
        const int nc = min(8, n_mt - c0);
        if (warp < nc) {
            const int row0 = (c0 + warp) * 16;
            float add[4], out[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) add[i] = pre(row0 + g + (i >> 1) * 8, 2 * t4 + (i & 1));
            S.stamp2();
            S.wait_full(S.cons + warp);
            S.stamp2();
            mma_mtile(S.slot_of(S.cons + warp), xs, DC_D * 2 + 64, 0, out, lane);
            S.stamp2();
            S.release(S.cons + warp, 8);
#pragma unroll
            for (int i = 0; i < 4; ++i)
                tr[(2 * t4 + (i & 1)) * 17 + g + (i >> 1) * 8] = fin(row0 + g + (i >> 1) * 8, 2 * t4 + (i & 1), out[i] + add[i]);
            __syncwarp();
            const float* src = tr + (lane >> 2) * 17 + (lane & 3) * 4;               // sequence lane/4, rows 4*(lane%4) ..
            put(row0 + (lane & 3) * 4, lane >> 2, make_float4(src[0], src[1], src[2], src[3]));
            __syncwarp();
            S.stamp2();
        }
        S.cons += nc;
        S.produce();
    }
}
// fc2 (K = 2048): tile (m-tile mt, k-chunk kq) = mt * 4 + kq; warp w takes tile w of every pass of 8 (2 m-tiles); the four
// k-chunks of an m-tile are summed through shared memory, thread (row = tid % 32, seq = tid / 32) finishes (a warp's
// stores are 32 consecutive floats of one sequence).
template <int CS, typename Pre, typename Post>
__device__ __forceinline__ void gemm_fc2(Stream<CS>& S, int n_mt, const unsigned char* xs, Pre pre, Post post) {
    const int warp = S.warp, lane = S.lane, g = lane >> 2, t4 = lane & 3;
    float* pb = reinterpret_cast<float*>(S.smem + OFF_PART);       // [2 m-tiles][4 kq][16 rows][8 seqs]
    for (int m0 = 0; m0 < n_mt; m0 += 2) {
        if (warp < 8) {
            const int mt = warp >> 2, kq = warp & 3;
            const int erow = m0 * 16 + lane, eseq = warp;                 // epilogue element of this thread (32 rows x 8 seqs)
            const float addend = pre(erow, eseq);
            float out[4];
            S.wait_full(S.cons + warp);
            mma_mtile(S.slot_of(S.cons + warp), xs, DC_FFN * 2 + 64, kq * DC_KC, out, lane);
            S.release(S.cons + warp, 8);
            if (m0 > 0) cbar();                                       // the previous pass's partials have been read
            *reinterpret_cast<float2*>(pb + ((mt * 4 + kq) * 16 + g) * 8 + 2 * t4) = make_float2(out[0], out[1]);
            *reinterpret_cast<float2*>(pb + ((mt * 4 + kq) * 16 + g + 8) * 8 + 2 * t4) = make_float2(out[2], out[3]);
            cbar();
            const float* pe = pb + ((lane >> 4) * 4 * 16 + (lane & 15)) * 8 + eseq;              // m-tile lane/16, row lane%16
            post(erow, eseq, (pe[0] + pe[128]) + (pe[256] + pe[384]) + addend);
        }
        S.cons += 8;
        S.produce();
    }
}

// ---- attention on the tensor cores, two passes over the (half-)unit's keys ----
// B operand of S = K q: column 0 = bf16(q'), column 1 = bf16(q' - bf16(q')), q' = q * head_dim^-1/2 * log2(e);
// q points at the 64 f32 of this head (staged in shared memory)
__device__ __forceinline__ void make_q_frags(const float* q, uint32_t (&qb)[4][2], int lane) {
    const int g = lane >> 2, t4 = lane & 3;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
        const float2 v0 = *reinterpret_cast<const float2*>(q + ks * 16 + 2 * t4);
        const float2 v1 = *reinterpret_cast<const float2*>(q + ks * 16 + 2 * t4 + 8);
        const float a0 = v0.x * DC_QSCALE, a1 = v0.y * DC_QSCALE, a2 = v1.x * DC_QSCALE, a3 = v1.y * DC_QSCALE;
        const uint32_t h0 = pack_bf16(a0, a1), h1 = pack_bf16(a2, a3);
        const uint32_t l0 = pack_bf16(a0 - bf16_round(a0), a1 - bf16_round(a1)), l1 = pack_bf16(a2 - bf16_round(a2), a3 - bf16_round(a3));
        qb[ks][0] = g == 0 ? h0 : (g == 1 ? l0 : 0u);
        qb[ks][1] = g == 0 ? h1 : (g == 1 ? l1 : 0u);
    }
}
// Attention of one query over n K tiles + n V tiles (tiles S.cons .. S.cons + 2n - 1 of the ring; keys key_first + 128 j + i,
// masked when >= n_valid).  Compute warps: warp w owns keys 16w..16w+15 of every tile.  Leaves the un-normalised
// output in s_acc[8 warps][64] (to be summed over the warps), the maximum (log2 domain) in red[16] and the 8 per-warp
// sums of p in red[8..15]; the caller synchronises (cbar) before reading them.  Uses sc = scores f32[768],
// p_hi / p_lo = bf16[768] each.
template <typename STR>
__device__ __forceinline__ void attn_two_pass(STR& S, int n, int key_first, int n_valid, const uint32_t (&qb)[4][2],
                                              float* sc, bf16* p_hi, bf16* p_lo, float* s_acc, float* red) {
    const int warp = S.warp, lane = S.lane;
    if (warp < 8 && n > 0) {
        const int mi = lane >> 3, rr = lane & 7, g = lane >> 2, t4 = lane & 3, tid = threadIdx.x;
        // ---- pass 1: S = K q for all K tiles (independent chains), scores to shared memory ----
        S.stamp2();
        S.wait_full_spread(S.cons, n);
        cbar();
        S.stamp2();
        {
            const uint32_t rowoff = (uint32_t)(16 * warp + (mi & 1) * 8 + rr) * 128u;
#pragma unroll
            for (int j = 0; j < DC_MAXHALF; ++j) {
                if (j < n) {
                    const uint32_t kt = smem_u32(S.slot_of(S.cons + j));
                    float c[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        uint32_t a0, a1, a2, a3;
                        ldsm_x4(kt + rowoff + ((uint32_t)((2 * ks + (mi >> 1)) ^ rr) << 4), a0, a1, a2, a3);
                        mma_bf16(c, a0, a1, a2, a3, qb[ks][0], qb[ks][1]);
                    }
                    S.release(S.cons + j, 1);
                    if (t4 == 0) {          // lanes with t4 == 0 hold keys 16w+g (c0+c1: q_hi + q_lo columns) and 16w+g+8 (c2+c3)
                        const int kl = j * DC_KEYS + 16 * warp + g, ka = key_first + kl;
                        sc[kl] = ka < n_valid ? c[0] + c[1] : -INFINITY;
                        sc[kl + 8] = ka + 8 < n_valid ? c[2] + c[3] : -INFINITY;
                    }
                }
            }
        }
        S.stamp2();
        cbar();
        S.stamp2();
        // ---- softmax over the n * 128 (<= 768) scores: max, p = 2^(s - max) as bf16 hi + lo, per-warp sums ----
        const int nk = n * DC_KEYS;
        float v[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) v[i] = tid + i * DC_CT < nk ? sc[tid + i * DC_CT] : -INFINITY;
        float mx = fmaxf(v[0], fmaxf(v[1], v[2]));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        if (lane == 0) red[warp] = mx;
        cbar();
        mx = fmaxf(fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3])), fmaxf(fmaxf(red[4], red[5]), fmaxf(red[6], red[7])));
        float lsum = 0.f;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            if (tid + i * DC_CT < nk) {
                const float p = ex2(v[i] - mx);
                const bf16 hi = __float2bfloat16(p);
                p_hi[tid + i * DC_CT] = hi;
                p_lo[tid + i * DC_CT] = __float2bfloat16(p - __bfloat162float(hi));
                lsum += p;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) lsum += __shfl_xor_sync(0xffffffffu, lsum, o);
        S.wait_full_spread(S.cons + n, n);  // the V tiles (published by the barrier below)
        cbar();                             // everyone has read red[0..7]; p arrays complete; V tiles visible
        if (lane == 0) { red[8 + warp] = lsum; if (warp == 0) red[16] = mx; }
        // ---- pass 2: O^T += V^T p over all V tiles; lane (g, t4): keys 2t4, 2t4+1 (b0) and 2t4+8, 2t4+9 (b1) of column g ----
        S.stamp2();
        S.stamp2();
        float o[4][4] = {};
        {
            const uint32_t rowoff = (uint32_t)(16 * warp + (mi >> 1) * 8 + rr) * 128u;
            const bf16* psel = g == 0 ? p_hi : p_lo;
#pragma unroll
            for (int j = 0; j < DC_MAXHALF; ++j) {
                if (j < n) {
                    const uint32_t vt = smem_u32(S.slot_of(S.cons + n + j));
                    const int kb = j * DC_KEYS + 16 * warp + 2 * t4;
                    const uint32_t b0 = g < 2 ? *reinterpret_cast<const uint32_t*>(psel + kb) : 0u;
                    const uint32_t b1 = g < 2 ? *reinterpret_cast<const uint32_t*>(psel + kb + 8) : 0u;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        uint32_t a0, a1, a2, a3;
                        ldsm_x4_t(vt + rowoff + ((uint32_t)((2 * i + (mi & 1)) ^ rr) << 4), a0, a1, a2, a3);
                        mma_bf16(o[i], a0, a1, a2, a3, b0, b1);
                    }
                    S.release(S.cons + n + j, 1);
                }
            }
        }
        if (t4 == 0) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                s_acc[warp * 64 + 16 * i + g] = o[i][0] + o[i][1];
                s_acc[warp * 64 + 16 * i + g + 8] = o[i][2] + o[i][3];
            }
        }
        S.stamp2();
    }
    S.cons += 2 * n;
    S.produce();
}

// Self-attention of ONE (sequence, head) unit by ONE warp: the cached keys are few (<= 448), so a warp runs the whole
// unit alone — 8 independent score chains per 128-key tile, an online softmax across tiles held in the warp, no CTA
// barrier — and the CTA's units run in parallel on different warps.  Tiles t0 .. t0+n-1 are the K tiles, t0+n .. t0+2n-1
// the V tiles (keys 0..s-1 valid).  `row` = staged q[64] | k[64] | v[64] of the new token (f32, shared memory); the new key
// and value are appended to the cache (bf16) and take part from registers.  Writes att[64] (bf16).
template <int CS>
__device__ __forceinline__ void self_attn_warp(Stream<CS>& S, int t0, int n, int s, const float* row, bf16* cache_row, bf16* att, int lane) {
    const int mi = lane >> 3, rr = lane & 7, g = lane >> 2, t4 = lane & 3;
    uint32_t qb[4][2];
    make_q_frags(row, qb, lane);
    // the new token: k, v rounded to bf16 (what later steps read back), score q . k_s in the log2 domain
    const float kn0 = bf16_round(row[64 + lane]), kn1 = bf16_round(row[64 + 32 + lane]);
    cache_row[lane] = __float2bfloat16(kn0);
    cache_row[32 + lane] = __float2bfloat16(kn1);
    cache_row[DC_D + lane] = __float2bfloat16(row[128 + lane]);
    cache_row[DC_D + 32 + lane] = __float2bfloat16(row[128 + 32 + lane]);
    float sc_new = row[lane] * DC_QSCALE * kn0 + row[32 + lane] * DC_QSCALE * kn1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sc_new += __shfl_xor_sync(0xffffffffu, sc_new, o);
    float m = -INFINITY, l = 0.f, o[4][4] = {};
    for (int j = 0; j < n; ++j) {
        S.wait_full(t0 + j);
        const uint32_t kt = smem_u32(S.slot_of(t0 + j));
        float c[8][4] = {};
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
#pragma unroll
            for (int mt = 0; mt < 8; ++mt) {
                uint32_t a0, a1, a2, a3;
                ldsm_x4(kt + (uint32_t)(16 * mt + (mi & 1) * 8 + rr) * 128u + ((uint32_t)((2 * ks + (mi >> 1)) ^ rr) << 4), a0, a1, a2, a3);
                mma_bf16(c[mt], a0, a1, a2, a3, qb[ks][0], qb[ks][1]);
            }
        S.release(t0 + j, 8);
        float sa[8], sb[8], mx = -INFINITY;
#pragma unroll
        for (int mt = 0; mt < 8; ++mt) {
            const int ka = j * DC_KEYS + 16 * mt + g;
            sa[mt] = (t4 == 0 && ka < s) ? c[mt][0] + c[mt][1] : -INFINITY;
            sb[mt] = (t4 == 0 && ka + 8 < s) ? c[mt][2] + c[mt][3] : -INFINITY;
            mx = fmaxf(mx, fmaxf(sa[mt], sb[mt]));
        }
#pragma unroll
        for (int of = 16; of > 0; of >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, of));
        const float mn = fmaxf(m, mx);                                  // finite: key 128 j is always valid (j < ceil(s / 128))
        const float scale = ex2(m - mn);
        m = mn;
        l *= scale;
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int k = 0; k < 4; ++k) o[i][k] *= scale;
        uint32_t b0[8], b1[8];
#pragma unroll
        for (int mt = 0; mt < 8; ++mt) {
            const float pa = ex2(sa[mt] - mn), pb = ex2(sb[mt] - mn);
            l += pa + pb;
            const float x0 = __shfl_sync(0xffffffffu, pa, 8 * t4), x1 = __shfl_sync(0xffffffffu, pa, 8 * t4 + 4);
            const float y0 = __shfl_sync(0xffffffffu, pb, 8 * t4), y1 = __shfl_sync(0xffffffffu, pb, 8 * t4 + 4);
            const uint32_t h0 = pack_bf16(x0, x1), h1 = pack_bf16(y0, y1);
            const uint32_t l0 = pack_bf16(x0 - bf16_round(x0), x1 - bf16_round(x1)), l1 = pack_bf16(y0 - bf16_round(y0), y1 - bf16_round(y1));
            b0[mt] = g == 0 ? h0 : (g == 1 ? l0 : 0u);
            b1[mt] = g == 0 ? h1 : (g == 1 ? l1 : 0u);
        }
        S.wait_full(t0 + n + j);
        const uint32_t vt = smem_u32(S.slot_of(t0 + n + j));
#pragma unroll
        for (int mt = 0; mt < 8; ++mt)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                uint32_t a0, a1, a2, a3;
                ldsm_x4_t(vt + (uint32_t)(16 * mt + (mi >> 1) * 8 + rr) * 128u + ((uint32_t)((2 * i + (mi & 1)) ^ rr) << 4), a0, a1, a2, a3);
                mma_bf16(o[i], a0, a1, a2, a3, b0[mt], b1[mt]);
            }
        S.release(t0 + n + j, 8);
    }
#pragma unroll
    for (int of = 16; of > 0; of >>= 1) l += __shfl_xor_sync(0xffffffffu, l, of);
    const float mt_ = fmaxf(m, sc_new), wo = ex2(m - mt_), wn = ex2(sc_new - mt_), inv = 1.0f / (l * wo + wn);
    if (t4 == 0) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int d0 = 16 * i + g, d1 = d0 + 8;
            att[d0] = __float2bfloat16(((o[i][0] + o[i][1]) * wo + bf16_round(row[128 + d0]) * wn) * inv);
            att[d1] = __float2bfloat16(((o[i][2] + o[i][3]) * wo + bf16_round(row[128 + d1]) * wn) * inv);
        }
    }
}

template <int CS>
__global__ void __launch_bounds__(DC_THREADS, 1)
dec_layers_kernel(const __grid_constant__ CUtensorMap tm_ckv, const __grid_constant__ CUtensorMap tm_skv, const DcArgs a) {
    extern __shared__ __align__(1024) unsigned char smem[];
    using St = Stream<CS>;
    St S(a, &tm_ckv, &tm_skv, smem);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool compute = warp < 8;                                  // warp 8 only issues tiles (S.produce)
    unsigned char* xs = smem + OFF_XS;
    // attention scratch inside OFF_PART (free between GEMM stages): scores | p_hi | p_lo | staged q (cross) or q|k|v (self)
    float* sc = reinterpret_cast<float*>(smem + OFF_PART);
    float* s_acc = reinterpret_cast<float*>(smem + OFF_ACC);        // [8][64]
    float* red = s_acc + 8 * 64;                                    // [64]
    float* s_bias = reinterpret_cast<float*>(smem + OFF_BIAS);      // [layer][qkv RQ | o RO | cq RO | co RO | fc1 RF | fc2 RO]
    float* x_own = reinterpret_cast<float*>(smem + OFF_XOWN);       // [RO rows][8 seqs]: this CTA's columns of the residual stream
    constexpr int NB = St::RQ + 4 * St::RO + St::RF, B_O = St::RQ, B_CQ = B_O + St::RO, B_CO = B_CQ + St::RO, B_F1 = B_CO + St::RO, B_F2 = B_F1 + St::RF;
    static_assert(St::RO <= 64, "scratch sizes");
    const bool hoist = NB <= DC_NBIAS && a.n_layers <= DC_BIAS_LAYERS;     // cluster of 16, <= 6 layers: biases live in shared memory
    const int rank = S.rank, b0 = S.b0, nb = S.nb;
    int n_stamp = 0;
    auto stamp = [&]() {
        if (a.prof != nullptr && blockIdx.x == 0 && tid == 0) a.prof[n_stamp] = clock64();
        ++n_stamp;
    };
    stamp();

    if (tid == 0) {
#pragma unroll
        for (int i = 0; i < DC_NSLOT; ++i) { mbar_init(S.full + i * 8, 1); mbar_init(S.empty + i * 8, 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    S.produce();                                    // q|k|v weights of layer 0: constant during a decode, in flight
    // this CTA's bias slices of every layer: constants too, fetched once instead of behind every stage's barrier
    for (int i = tid; hoist && i < a.n_layers * NB; i += DC_THREADS) {
        const int l = i / NB, j = i - l * NB;
        const DcLayer& L = a.layers[l];
        float v;
        if (j < B_O) v = L.bqkv[rank * St::RQ + j];
        else if (j < B_CQ) v = L.bo[rank * St::RO + j - B_O];
        else if (j < B_CO) v = L.bcq[rank * St::RO + j - B_CQ];
        else if (j < B_F1) v = L.bco[rank * St::RO + j - B_CO];
        else if (j < B_F2) v = L.bfc1[rank * St::RF + j - B_F1];
        else v = L.bfc2[rank * St::RO + j - B_F2];
        s_bias[i] = v;
    }
    float4 gw[4], gb[4];
    ln_params(a.layers[0].ln1w, a.layers[0].ln1b, gw, gb, lane);
    asm volatile("griddepcontrol.wait;" ::: "memory");           // before the predecessor kernel (arg-max of the last step) is done
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const int s = a.state[0], prompt_len = a.state[1];
    S.build_table(s);
    __syncthreads();                                // last CTA-wide barrier: from here on the producer warp only joins cluster barriers
    S.produce();
    stamp();

    for (int l = 0; l < a.n_layers; ++l) {
        const DcLayer& L = a.layers[l];
        const float* bias = s_bias + l * NB;
        const float* b_qkv = hoist ? bias : L.bqkv + rank * St::RQ;
        const float* b_o = hoist ? bias + B_O : L.bo + rank * St::RO;
        const float* b_cq = hoist ? bias + B_CQ : L.bcq + rank * St::RO;
        const float* b_co = hoist ? bias + B_CO : L.bco + rank * St::RO;
        const float* b_f1 = hoist ? bias + B_F1 : L.bfc1 + rank * St::RF;
        const float* b_f2 = hoist ? bias + B_F2 : L.bfc2 + rank * St::RO;
        bf16* skv = a.self_kv + (size_t)l * a.Bmax * a.T_max * 2 * DC_D;

        // ---------------- stage 1: LN1 + fused q|k|v projection ----------------
        if (compute) {
            if (l == 0) {
                // token + position embedding, x[b] = E[tok] + P[s] (whole row for the LayerNorm; this CTA's column slice
                // goes to the residual stream, kept in shared memory for its owner and in global memory for the others)
                unsigned char* xrow = xs + warp * DC_WROW;
                if (warp < nb) {
                    const int b = b0 + warp;
                    const int tok = s < prompt_len ? a.prompt[s] : a.cur_tok[b];
                    float4 v[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int c = i * 128 + lane * 4;
                        const uint2 e = *reinterpret_cast<const uint2*>(a.E + (size_t)tok * DC_D + c);
                        const float4 p = *reinterpret_cast<const float4*>(a.P + (size_t)s * DC_D + c);
                        v[i] = make_float4(__uint_as_float(e.x << 16) + p.x, __uint_as_float(e.x & 0xffff0000u) + p.y,
                                           __uint_as_float(e.y << 16) + p.z, __uint_as_float(e.y & 0xffff0000u) + p.w);
                        if (c >= rank * St::RO && c < (rank + 1) * St::RO) {
                            const int r = c - rank * St::RO;
                            x_own[(r + 0) * 8 + warp] = v[i].x; x_own[(r + 1) * 8 + warp] = v[i].y;
                            x_own[(r + 2) * 8 + warp] = v[i].z; x_own[(r + 3) * 8 + warp] = v[i].w;
                        }
                    }
                    ln_pack_row(v, gw, gb, xrow, lane);
                } else {
                    zero_row(xrow, DC_D, lane);
                }
            } else {
                stage_ln(a.x, b0, nb, gw, gb, xs);
            }
            cbar();
        }
        gemm_rows<CS>(S, St::NQ, xs,
            [&](int row, int) { return b_qkv[row]; },
            [&](int, int, float v) { return v; },
            [&](int row4, int seq, float4 v) {
                if (seq < nb) *reinterpret_cast<float4*>(a.qkv + (size_t)(b0 + seq) * 3 * DC_D + rank * St::RQ + row4) = v;
            });
        cluster_arrive();
        cluster_wait();
        stamp();

        // ---------------- stage 2: causal self-attention of the new token over the cached keys 0..s ----------------
        // keys 0..s-1 come from the cache through the ring (written by earlier steps), the new key from registers.
        // One L2 round trip fetches the q|k|v head slices of all this CTA's units; then one warp per unit.
        {
            float* stg = sc;                                        // staged q|k|v of <= 8 units
            if (compute) {
                for (int idx = tid; idx < S.n_self_units * 192; idx += DC_CT) {
                    const int u = idx / 192, e = idx - u * 192, unit = rank + u * CS;
                    stg[idx] = __ldcg(a.qkv + (size_t)(b0 + unit / DC_H) * 3 * DC_D + (e >> 6) * DC_D + (unit % DC_H) * DC_HD + (e & 63));
                }
                cbar();
                if (warp < S.n_self_units) {
                    const int unit = rank + warp * CS, b = b0 + unit / DC_H, h = unit % DC_H;
                    self_attn_warp<CS>(S, S.cons + warp * 2 * S.nst, S.nst, s, stg + warp * 192,
                                       skv + ((size_t)b * a.T_max + s) * 2 * DC_D + h * DC_HD, a.att + (size_t)b * DC_D + h * DC_HD, lane);
                }
            }
            S.cons += S.n_self_units * 2 * S.nst;
            S.produce();
        }
        cluster_arrive();
        cluster_wait();
        stamp();

        // ---------------- stage 3: self-attention out-projection + residual ----------------
        if (compute) {
            stage_bf16(a.att, DC_D, b0, nb, xs);
            cbar();
        }
        gemm_rows<CS>(S, St::NO, xs,
            [&](int row, int seq) { return b_o[row] + x_own[row * 8 + seq]; },
            [&](int row, int seq, float v) { x_own[row * 8 + seq] = v; return v; },
            [&](int row4, int seq, float4 v) {
                if (seq < nb) *reinterpret_cast<float4*>(a.x + (size_t)(b0 + seq) * DC_D + rank * St::RO + row4) = v;
            });
        cluster_arrive();
        ln_params(L.ln2w, L.ln2b, gw, gb, lane);
        cluster_wait();
        stamp();

        // ---------------- stage 4: LN2 + cross-attention query projection ----------------
        if (compute) {
            stage_ln(a.x, b0, nb, gw, gb, xs);
            cbar();
        }
        gemm_rows<CS>(S, St::NO, xs,
            [&](int row, int) { return b_cq[row]; },
            [&](int, int, float v) { return v; },
            [&](int row4, int seq, float4 v) {
                if (seq < nb) *reinterpret_cast<float4*>(a.q + (size_t)(b0 + seq) * DC_D + rank * St::RO + row4) = v;
            });
        cluster_arrive();
        cluster_wait();
        stamp();

        // ---------------- stage 5: cross-attention over the cached encoder K/V (the HBM stream) ----------------
        // this CTA's share: nhu half-units (one head of one sequence, first or second half of the keys)
        {
            bf16* p_hi = reinterpret_cast<bf16*>(sc + DC_MAXHALF * DC_KEYS);     // cross: 768 scores | p_hi | p_lo | staged q of <= 8 half-units
            bf16* p_lo = p_hi + DC_MAXHALF * DC_KEYS;
            float* stg = sc + 2 * DC_MAXHALF * DC_KEYS;
            S.n2 = (l == 1 && a.prof_stage == 5) ? 256 : 1 << 30;
            S.stamp2();
            uint32_t* qfr = reinterpret_cast<uint32_t*>(xs);        // [half-unit][ks 4][reg 2][hi/lo 2][t4 4] packed B fragments (xs is free here)
            if (compute) {
                for (int idx = tid; idx < S.nhu * DC_HD; idx += DC_CT) {           // their q slices: one L2 round trip
                    const int unit = (rank * S.nhu + (idx >> 6)) >> 1;
                    stg[idx] = __ldcg(a.q + (size_t)(b0 + unit / DC_H) * DC_D + (unit % DC_H) * DC_HD + (idx & 63));
                }
                cbar();
                for (int idx = tid; idx < S.nhu * 64; idx += DC_CT) {
                    const int hi = idx >> 6, w = idx & 63, ks = w >> 4, reg = (w >> 3) & 1, lo = (w >> 2) & 1, t4 = w & 3;
                    const float* q = stg + hi * DC_HD + ks * 16 + 2 * t4 + reg * 8;
                    const float a0 = q[0] * DC_QSCALE, a1 = q[1] * DC_QSCALE;
                    qfr[idx] = lo ? pack_bf16(a0 - bf16_round(a0), a1 - bf16_round(a1)) : pack_bf16(a0, a1);
                }
                cbar();
            }
            for (int hi = 0; hi < S.nhu; ++hi) {
                const int hu = rank * S.nhu + hi, unit = hu >> 1, half = hu & 1, b = b0 + unit / DC_H, h = unit % DC_H;
                uint32_t qb[4][2];
                S.stamp2();
                if (compute) {
                    const int g = lane >> 2, t4 = lane & 3;
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        qb[ks][0] = g < 2 ? qfr[hi * 64 + ks * 16 + g * 4 + t4] : 0u;
                        qb[ks][1] = g < 2 ? qfr[hi * 64 + ks * 16 + 8 + g * 4 + t4] : 0u;
                    }
                }
                attn_two_pass(S, S.tph, half * S.tph * DC_KEYS, a.Tk, qb, sc, p_hi, p_lo, s_acc, red);
                S.stamp2();
                if (compute) {
                    cbar();
                    if (tid < 64) {
                        float acc = 0.f, lt = 0.f;
#pragma unroll
                        for (int w = 0; w < 8; ++w) { acc += s_acc[w * 64 + tid]; lt += red[8 + w]; }
                        float* rec = a.xpart + (((size_t)b * DC_H + h) * 2 + half) * 66;
                        rec[tid] = acc;
                        if (tid == 0) { rec[64] = red[16]; rec[65] = lt; }
                    }
                    cbar();                                         // scratch is reused by the next half-unit
                }
            }
            S.stamp2();
            S.n2 = 1 << 30;
        }
        cluster_arrive();
        cluster_wait();
        stamp();

        // ---------------- stage 6: cross-attention out-projection + residual ----------------
        // staging merges the two key halves of every (sequence, head) unit
        if (compute) {
            for (int idx = tid; idx < DC_SEQ * DC_D / 2; idx += DC_CT) {
                const int row = idx / (DC_D / 2), c = (idx - row * (DC_D / 2)) * 2;        // two neighbouring dims
                uint32_t pk = 0u;
                if (row < nb) {
                    const int h = c / DC_HD;
                    const float* r0 = a.xpart + (((size_t)(b0 + row) * DC_H + h) * 2) * 66;
                    const float* r1 = r0 + 66;
                    const float2 a0 = __ldcg(reinterpret_cast<const float2*>(r0 + (c - h * DC_HD)));
                    const float2 a1 = __ldcg(reinterpret_cast<const float2*>(r1 + (c - h * DC_HD)));
                    const float m0 = __ldcg(r0 + 64), l0 = __ldcg(r0 + 65), m1 = __ldcg(r1 + 64), l1 = __ldcg(r1 + 65);
                    const float mt = fmaxf(m0, m1), w0 = ex2(m0 - mt), w1 = ex2(m1 - mt);
                    const float lt = l0 * w0 + l1 * w1;
                    pk = pack_bf16((a0.x * w0 + a1.x * w1) / lt, (a0.y * w0 + a1.y * w1) / lt);
                }
                *reinterpret_cast<uint32_t*>(xs + row * DC_WROW + c * 2) = pk;
            }
            cbar();
        }
        gemm_rows<CS>(S, St::NO, xs,
            [&](int row, int seq) { return b_co[row] + x_own[row * 8 + seq]; },
            [&](int row, int seq, float v) { x_own[row * 8 + seq] = v; return v; },
            [&](int row4, int seq, float4 v) {
                if (seq < nb) *reinterpret_cast<float4*>(a.x + (size_t)(b0 + seq) * DC_D + rank * St::RO + row4) = v;
            });
        S.n2 = (l == 1 && a.prof_stage == 7) ? 256 : 1 << 30;
        S.stamp2();
        cluster_arrive();
        S.stamp2();
        ln_params(L.ln3w, L.ln3b, gw, gb, lane);
        cluster_wait();
        S.stamp2();
        stamp();

        // ---------------- stage 7: LN3 + fc1 + GELU ----------------
        if (compute) {
            stage_ln(a.x, b0, nb, gw, gb, xs);
            S.stamp2();
            cbar();
            S.stamp2();
        }
        gemm_rows<CS>(S, St::NF1, xs,
            [&](int row, int) { return b_f1[row]; },
            [&](int, int, float v) { return gelu_erf(v); },
            [&](int row4, int seq, float4 v) {
                if (seq < nb)
                    *reinterpret_cast<uint2*>(a.ffn + (size_t)(b0 + seq) * DC_FFN + rank * St::RF + row4) = make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
            });
        S.stamp2();
        cluster_arrive();
        S.stamp2();
        cluster_wait();
        S.stamp2();
        if (a.prof_stage == 7 && a.prof != nullptr) {       // experiment: a second barrier with no stores in between (fixed cost of release / acquire)
            cluster_arrive();
            S.stamp2();
            cluster_wait();
            S.stamp2();
            asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
            S.stamp2();
            asm volatile("barrier.cluster.wait.aligned;" ::: "memory");
            S.stamp2();
        }
        S.n2 = 1 << 30;
        stamp();

        // ---------------- stage 8: fc2 + residual ----------------
        if (compute) {
            stage_bf16(a.ffn, DC_FFN, b0, nb, xs);
            cbar();
        }
        gemm_fc2<CS>(S, St::NO, xs,
            [&](int row, int seq) { return b_f2[row] + x_own[row * 8 + seq]; },
            [&](int row, int seq, float v) {
                x_own[row * 8 + seq] = v;
                if (seq < nb) a.x[(size_t)(b0 + seq) * DC_D + rank * St::RO + row] = v;
            });
        if (l + 1 < a.n_layers) {
            cluster_arrive();
            ln_params(a.layers[l + 1].ln1w, a.layers[l + 1].ln1b, gw, gb, lane);
            cluster_wait();
        }
        stamp();
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Stand-alone cross-attention of one decoder layer for the per-kernel decode path (K3e, the dominant HBM stream of a
// step: 2 x Tk x d bf16 per sequence): the same tile machinery as above — a producer warp streams 128-key K and V tiles
// (3-D tensor-map tiles, 128B swizzle) through a 10-slot ring, 8 compute warps run the two-pass tensor-core attention per
// key half — as a persistent kernel, one CTA per SM walking (sequence, head) units.  Replaces cross_attn_kernel<bf16>
// (lane-group LDG streaming: 20.6 us = 0.73 of the measured HBM peak at B = 32, 5.7 us of it launch / ramp / merge tail):
// the ring is filled before the PDL wait and keeps ~160 KB per SM in flight across unit boundaries.
constexpr int XA_NSLOT = 10;
constexpr int XA_OFF_SC = XA_NSLOT * DC_TILE;                      // scores f32[768] | p_hi bf16[768] | p_lo bf16[768]
constexpr int XA_OFF_QFR = XA_OFF_SC + 6144;                       // packed q fragments [64] u32 + staged q f32[64]
constexpr int XA_OFF_ACC = XA_OFF_QFR + 512;                       // [8][64] + red[64]
constexpr int XA_OFF_BAR = XA_OFF_ACC + (8 * 64 + 64) * 4;
constexpr int XA_SMEM = XA_OFF_BAR + 2 * XA_NSLOT * 8 + 16;

struct XaArgs {
    const float* q;                 // [B][d] f32
    float* out;                     // [B][d] f32
    int H, B, d, Tk, zbase;         // zbase = layer * Bmax (third tensor-map coordinate of sequence 0)
};
struct XStream {                    // the slice of Stream<> that attn_two_pass uses
    const XaArgs& a;
    const CUtensorMap* tm;
    unsigned char* smem;
    uint32_t full, empty;
    int warp, lane, cons = 0, issued = 0, total, tph, nkt, units_mine;
    __device__ __forceinline__ XStream(const XaArgs& a_, const CUtensorMap* tm_, unsigned char* sm) : a(a_), tm(tm_), smem(sm) {
        warp = threadIdx.x >> 5; lane = threadIdx.x & 31;
        full = smem_u32(sm + XA_OFF_BAR); empty = full + XA_NSLOT * 8;
        nkt = (a.Tk + DC_KEYS - 1) / DC_KEYS; tph = nkt / 2;
        const int units = a.B * a.H;
        units_mine = (int)blockIdx.x < units ? (units - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
        total = units_mine * 2 * nkt;
    }
    __device__ __forceinline__ void stamp2() {}
    __device__ __forceinline__ unsigned char* slot_of(int t) const { return smem + (t % XA_NSLOT) * DC_TILE; }
    __device__ __forceinline__ void wait_full(int t) { mbar_wait(full + (t % XA_NSLOT) * 8, (uint32_t)(t / XA_NSLOT) & 1u); }
    __device__ __forceinline__ void wait_full_spread(int t, int n) { if (warp < n) wait_full(t + warp); }
    __device__ __forceinline__ void release(int t, int n) {
        __syncwarp();
        if (lane == 0) mbar_arrive_n(empty + (t % XA_NSLOT) * 8, (uint32_t)n);
    }
    // tile t of this CTA: unit t / (2 nkt); inside the unit: half 0 = K tiles 0..tph-1 then V tiles 0..tph-1, half 1 the rest
    __device__ __forceinline__ void produce() {
        if (warp != 8) return;
        const int want = min(cons + XA_NSLOT, total);
        while (issued < want) {
            if (lane == 0) {
                const int slot = issued % XA_NSLOT;
                if (issued >= XA_NSLOT) mbar_wait(empty + slot * 8, (uint32_t)(issued / XA_NSLOT - 1) & 1u);
                const int ui = issued / (2 * nkt), r = issued - ui * 2 * nkt, half = r / (2 * tph), j = r - half * 2 * tph;
                const int unit = (int)blockIdx.x + ui * (int)gridDim.x, b = unit / a.H, h = unit - b * a.H;
                const int v = j >= tph, key0 = (half * tph + (j - v * tph)) * DC_KEYS;
                mbar_expect_tx(full + slot * 8, DC_TILE);
                tma_tile_3d(tm, full + slot * 8, smem + slot * DC_TILE, v * a.d + h * DC_HD, key0, a.zbase + b);
            }
            ++issued;
        }
    }
};

__global__ void __launch_bounds__(DC_THREADS, 1)
cross_attn_tc_kernel(const __grid_constant__ CUtensorMap tm_ckv, const XaArgs a) {
    extern __shared__ __align__(1024) unsigned char smem[];
    XStream S(a, &tm_ckv, smem);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool compute = warp < 8;
    float* sc = reinterpret_cast<float*>(smem + XA_OFF_SC);
    bf16* p_hi = reinterpret_cast<bf16*>(sc + DC_MAXHALF * DC_KEYS);
    bf16* p_lo = p_hi + DC_MAXHALF * DC_KEYS;
    uint32_t* qfr = reinterpret_cast<uint32_t*>(smem + XA_OFF_QFR);
    float* qst = reinterpret_cast<float*>(qfr + 64);
    float* s_acc = reinterpret_cast<float*>(smem + XA_OFF_ACC);
    float* red = s_acc + 8 * 64;
    if (tid == 0) {
#pragma unroll
        for (int i = 0; i < XA_NSLOT; ++i) { mbar_init(S.full + i * 8, 1); mbar_init(S.empty + i * 8, 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    S.produce();                                    // the cached K/V do not depend on the predecessor kernel: the ring fills now
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    for (int ui = 0; ui < S.units_mine; ++ui) {
        const int unit = (int)blockIdx.x + ui * (int)gridDim.x, b = unit / a.H, h = unit - b * a.H;
        uint32_t qb[4][2];
        if (compute) {
            if (tid < 64) qst[tid] = a.q[(size_t)b * a.d + h * DC_HD + tid];
            cbar();
            if (tid < 64) {
                const int w = tid, ks = w >> 4, reg = (w >> 3) & 1, lo = (w >> 2) & 1, t4 = w & 3;
                const float* q = qst + ks * 16 + 2 * t4 + reg * 8;
                const float a0 = q[0] * DC_QSCALE, a1 = q[1] * DC_QSCALE;
                qfr[w] = lo ? pack_bf16(a0 - bf16_round(a0), a1 - bf16_round(a1)) : pack_bf16(a0, a1);
            }
            cbar();
            const int g = lane >> 2, t4 = lane & 3;
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                qb[ks][0] = g < 2 ? qfr[ks * 16 + g * 4 + t4] : 0u;
                qb[ks][1] = g < 2 ? qfr[ks * 16 + 8 + g * 4 + t4] : 0u;
            }
        }
        float acc0 = 0.f, m0 = -INFINITY, l0 = 0.f;                  // first half's state (threads 0..63: one head dim each)
        for (int half = 0; half < 2; ++half) {
            const int n = half == 0 ? S.tph : S.nkt - S.tph;
            attn_two_pass(S, n, half * S.tph * DC_KEYS, a.Tk, qb, sc, p_hi, p_lo, s_acc, red);
            if (compute) {
                cbar();
                if (tid < 64) {
                    float acc = 0.f, lt = 0.f;
#pragma unroll
                    for (int w = 0; w < 8; ++w) { acc += s_acc[w * 64 + tid]; lt += red[8 + w]; }
                    const float mh = red[16];
                    if (half == 0) { acc0 = acc; m0 = mh; l0 = lt; }
                    else {
                        const float mt = fmaxf(m0, mh), w0 = ex2(m0 - mt), w1 = ex2(mh - mt);
                        a.out[(size_t)b * a.d + h * DC_HD + tid] = (acc0 * w0 + acc * w1) / (l0 * w0 + lt * w1);
                    }
                }
                cbar();                                             // scratch is reused by the next half / unit
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Final LayerNorm + tied vocabulary projection + masked arg-max + token bookkeeping of one decode step in ONE launch
// (K3h; argmax_last_dim_raw, /root/reference/src/main.rs:709-735, and the loop control of :777-783 / :816-826).
// The [V][512] embedding is streamed once per step (53 MB of the step's weights): every CTA walks 32-row chunks of a
// pitched copy through a TMA ring (producer warp + full/empty mbarriers), 8 MMA warps = 2 row tiles x 4 sequence tiles
// (up to 32 sequences) with the whole K per warp, arg-max taken straight from the accumulator fragments.  The last CTA
// to finish (arrival counter) merges the per-CTA partials, writes the tokens and advances the step counter.
constexpr int DV_THREADS = 256, DV_NSLOT = 5, DV_SEQ = 32, DV_CROWS = 32, DV_SLOT = DV_CROWS * DC_WROW;
constexpr int DV_OFF_XS = DV_NSLOT * DV_SLOT;
constexpr int DV_OFF_RED = DV_OFF_XS + DV_SEQ * DC_WROW;           // [8 warps][8 seqs] val + idx
constexpr int DV_OFF_BAR = DV_OFF_RED + 8 * 8 * 8;
constexpr int DV_SMEM = DV_OFF_BAR + DV_NSLOT * 8 + 16;

struct DvArgs {
    int* state;                     // [0] = position s (advanced by the last CTA), [1] = prompt_len
    const float* x;                 // [B][d] output of the last decoder layer
    const float *lnw, *lnb;
    const bf16* E;                  // [V][d] tied embedding
    int V, B, n_chunks;
    const unsigned *sup_base, *sup_first;
    float* logits;                  // optional [B][V]
    float* pval; int* pidx;         // [grid][32] per-CTA partial arg-max
    unsigned int* counter;          // arrival counter (zero between launches)
    const int* forced; int max_new, eot, T_total;
    int *tokens, *lens, *finished, *cur_tok;
    int* unfinished_out;            // mapped host int: sequences still running after this step
};

__device__ __forceinline__ bool better(float v1, int i1, float v2, int i2) { return v1 > v2 || (v1 == v2 && i1 < i2); }

__global__ void __launch_bounds__(DV_THREADS, 1)
dec_vocab_kernel(const DvArgs a) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ int s_last;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    unsigned char* xs = smem + DV_OFF_XS;
    float* red_v = reinterpret_cast<float*>(smem + DV_OFF_RED);
    int* red_i = reinterpret_cast<int*>(red_v + 64);
    const uint32_t full = smem_u32(smem + DV_OFF_BAR);
    const int my_chunks = a.n_chunks > (int)blockIdx.x ? (a.n_chunks - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    if (tid == 0) {
#pragma unroll
        for (int i = 0; i < DV_NSLOT; ++i) mbar_init(full + i * 8, DV_THREADS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // chunk i of this CTA (32 embedding rows = 32 KB contiguous) into slot i % NSLOT: 8 LDGSTS per thread, rows padded to
    // 1088 B in shared memory, rows past the vocabulary zero-filled
    const uint32_t my_dst = (uint32_t)((tid >> 6) * DC_WROW + (tid & 63) * 16);      // copy j: row tid/64 + 4j, 16-byte piece tid%64
    const bf16* my_src = a.E + (size_t)(tid >> 6) * DC_D + (tid & 63) * 8;
    auto issue = [&](int i) {
        const int slot = i % DV_NSLOT, n0 = (int)(blockIdx.x + (size_t)i * gridDim.x) * DV_CROWS;
        const uint32_t d0 = smem_u32(smem + slot * DV_SLOT) + my_dst;
        const bf16* src = my_src + (size_t)n0 * DC_D;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const bool ok = n0 + (tid >> 6) + 4 * j < a.V;
            cp_async16(d0 + j * 4 * DC_WROW, ok ? src + (size_t)j * 4 * DC_D : a.E, ok ? 16u : 0u);
        }
        cp_async_arrive(full + slot * 8);
    };
    for (int i = 0; i < min(DV_NSLOT, my_chunks); ++i) issue(i);     // the embedding is constant: the ring fills while the
    asm volatile("griddepcontrol.wait;" ::: "memory");               // layer kernel still runs
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    float bestv[2] = {-INFINITY, -INFINITY};
    int besti[2] = {0x7fffffff, 0x7fffffff};
    {
        // final LayerNorm of every sequence -> bf16 rows (warp w: rows w, w+8, w+16, w+24)
        float4 gw[4], gb[4];
        ln_params(a.lnw, a.lnb, gw, gb, lane);
        for (int r = warp; r < DV_SEQ; r += 8) {
            unsigned char* xrow = xs + r * DC_WROW;
            if (r < a.B) {
                float4 v[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) v[i] = *reinterpret_cast<const float4*>(a.x + (size_t)r * DC_D + i * 128 + lane * 4);
                ln_pack_row(v, gw, gb, xrow, lane);
            } else {
                zero_row(xrow, DC_D, lane);
            }
        }
        __syncthreads();
        const int s = a.state[0], gi = s - (a.state[1] - 1);
        const unsigned* sup = gi == 0 ? a.sup_first : a.sup_base;
        // a chunk is 32 aligned vocabulary rows = ONE word of the suppress bitmap: fetch this CTA's words up front (a
        // dependent global load per chunk sat on the critical path of every chunk)
        unsigned supw[12];
#pragma unroll
        for (int i = 0; i < 12; ++i) supw[i] = i < my_chunks ? sup[blockIdx.x + i * gridDim.x] : 0xffffffffu;
        const int g = lane >> 2, t4 = lane & 3, mt = warp & 1, nt = warp >> 1;
        const int seq0 = nt * 8 + 2 * t4;
#pragma unroll 1
        for (int i = 0; i < my_chunks; ++i) {
            const int slot = i % DV_NSLOT;
            unsigned sw = 0xffffffffu;
#pragma unroll
            for (int q = 0; q < 12; ++q) sw = q == i ? supw[q] : sw;
            if (i >= 12) sw = sup[blockIdx.x + i * gridDim.x];
            mbar_wait(full + slot * 8, (uint32_t)(i / DV_NSLOT) & 1u);
            const unsigned char* a0 = smem + slot * DV_SLOT + (mt * 16 + g) * DC_WROW + 8 * t4 * 2;
            const unsigned char* bx = xs + (nt * 8 + g) * DC_WROW + 8 * t4 * 2;
            float ac[4][4] = {};                                     // four interleaved accumulators: 8 dependent mma each
#pragma unroll
            for (int c = 0; c < DC_KC / 32; ++c) {
                const uint4 wa = *reinterpret_cast<const uint4*>(a0 + c * 64);
                const uint4 wb = *reinterpret_cast<const uint4*>(a0 + 8 * DC_WROW + c * 64);
                const uint4 xb = *reinterpret_cast<const uint4*>(bx + c * 64);
                mma_bf16(ac[(2 * c) & 3], wa.x, wb.x, wa.y, wb.y, xb.x, xb.y);
                mma_bf16(ac[(2 * c + 1) & 3], wa.z, wb.z, wa.w, wb.w, xb.z, xb.w);
            }
            float acc[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[q] = (ac[0][q] + ac[1][q]) + (ac[2][q] + ac[3][q]);
            __syncthreads();                                         // every warp is done with the slot: refill it
            if (i + DV_NSLOT < my_chunks) issue(i + DV_NSLOT);
            const int n0 = (int)(blockIdx.x + (size_t)i * gridDim.x) * DV_CROWS + mt * 16 + g;
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                const int n = n0 + hh * 8;
                if (n < a.V) {
                    const bool ok = !((sw >> (n & 31)) & 1u);
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        const float v = acc[hh * 2 + j];
                        if (a.logits && seq0 + j < a.B) a.logits[(size_t)(seq0 + j) * a.V + n] = v;
                        if (ok && better(v, n, bestv[j], besti[j])) { bestv[j] = v; besti[j] = n; }    // strict '>', lowest index, NaN never
                    }
                }
            }
        }
        // lanes sharing a sequence pair (same t4) over the 8 row lanes g
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int o = 4; o < 32; o <<= 1) {
                const float ov = __shfl_xor_sync(0xffffffffu, bestv[j], o);
                const int oi = __shfl_xor_sync(0xffffffffu, besti[j], o);
                if (better(ov, oi, bestv[j], besti[j])) { bestv[j] = ov; besti[j] = oi; }
            }
        if (g == 0) {
#pragma unroll
            for (int j = 0; j < 2; ++j) { red_v[warp * 8 + 2 * t4 + j] = bestv[j]; red_i[warp * 8 + 2 * t4 + j] = besti[j]; }
        }
        __syncthreads();
        if (tid < DV_SEQ) {          // sequence tid: row tiles mt = 0, 1 of sequence tile nt = tid / 8
            const int w0 = (tid >> 3) * 2, c = tid & 7;
            float bv = red_v[w0 * 8 + c];
            int bi = red_i[w0 * 8 + c];
            if (better(red_v[(w0 + 1) * 8 + c], red_i[(w0 + 1) * 8 + c], bv, bi)) { bv = red_v[(w0 + 1) * 8 + c]; bi = red_i[(w0 + 1) * 8 + c]; }
            a.pval[blockIdx.x * DV_SEQ + tid] = bv;
            a.pidx[blockIdx.x * DV_SEQ + tid] = bi;
        }
    }
    // ---- last CTA to arrive: merge the partials, token bookkeeping, advance the step ----
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = (atomicAdd(a.counter, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    const int s = a.state[0], prompt_len = a.state[1], gi = s - (prompt_len - 1);
    {
        for (int b = warp; b < a.B; b += 8) {
            float bv = -INFINITY;
            int bi = 0x7fffffff;
            for (int p = lane; p < (int)gridDim.x; p += 32) {
                const float v = __ldcg(a.pval + p * DV_SEQ + b);
                const int i = __ldcg(a.pidx + p * DV_SEQ + b);
                if (better(v, i, bv, bi)) { bv = v; bi = i; }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
            }
            if (lane == 0) {
                const int tok = (bi == 0x7fffffff) ? 0 : bi;          // nothing beat -inf -> index 0
                if (!a.finished[b]) {
                    a.tokens[(size_t)b * a.T_total + prompt_len + gi] = tok;
                    a.lens[b] = prompt_len + gi + 1;
                    if (tok == a.eot) a.finished[b] = 1;               // main.rs:781-783, 820-822
                }
                a.cur_tok[b] = a.forced ? a.forced[(size_t)b * a.max_new + gi] : tok;
            }
        }
    }
    __syncthreads();
    if (tid == 0) {
        int n = 0;
        for (int b = 0; b < a.B; ++b) n += a.finished[b] ? 0 : 1;
        if (a.unfinished_out) *a.unfinished_out = n;
        a.state[0] = s + 1;
        *a.counter = 0u;
    }
}

// weights [N][K] bf16 -> chunk images [N/16][K/512][16][544]
__global__ void dc_pack_kernel(const bf16* __restrict__ src, bf16* __restrict__ dst, int N, int K) {
    const int KQ = K / DC_KC;
    const size_t n8 = (size_t)N * K / 8;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n8; i += (size_t)gridDim.x * blockDim.x) {
        const int n = (int)(i / (K / 8)), k = (int)(i % (K / 8)) * 8;
        const size_t chunk = (size_t)(n / DC_CROWS) * KQ + k / DC_KC;
        *reinterpret_cast<uint4*>(dst + (chunk * DC_CROWS + n % DC_CROWS) * DC_WPITCH + k % DC_KC) =
            *reinterpret_cast<const uint4*>(src + (size_t)n * K + k);
    }
}

// same footprint as dec_layers_kernel (threads, dynamic shared memory): how many clusters of each size the part can
// hold at once (depends on how many SMs each GPC has left after yield harvesting)
__global__ void __launch_bounds__(DC_THREADS, 1) dc_occupancy_probe_kernel(int* out) {
    extern __shared__ __align__(16) unsigned char sm[];
    if (out && threadIdx.x == 0) out[blockIdx.x] = sm[0];
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int CS>
cudaError_t launch_layers(cudaStream_t st, bool pdl, int n_clusters, const CUtensorMap& tmc, const CUtensorMap& tms, const DcArgs& a) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(n_clusters * CS);
    cfg.blockDim = dim3(DC_THREADS);
    cfg.dynamicSmemBytes = DC_SMEM;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    int na = 0;
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = CS; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1;
    ++na;
    if (pdl) {
        attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    cfg.attrs = attr;
    cfg.numAttrs = na;
    return cudaLaunchKernelEx(&cfg, dec_layers_kernel<CS>, tmc, tms, a);
}

template <typename K>
int max_active_clusters_of(K kernel, int cs) {
    if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DC_SMEM) != cudaSuccess) { cudaGetLastError(); return 0; }
    if (cs > 8 && cudaFuncSetAttribute(kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) { cudaGetLastError(); return 0; }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(cs * 16); cfg.blockDim = dim3(DC_THREADS); cfg.dynamicSmemBytes = DC_SMEM;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cs; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, kernel, &cfg) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

}  // namespace

// ---- host side ----
struct DecCluster {
    CUtensorMap tm_ckv, tm_skv;
    DevBuf<unsigned char> layers;       // DcLayer[n_layers]
    DevBuf<unsigned char> wimg;         // pitched weight chunk images
    DevBuf<float> pval;
    DevBuf<int> pidx;
    DevBuf<unsigned int> counter;
    int n_vchunks = 0;
    // which of the optional kernels this context uses: read from the environment when the context is created
    // (WB_DEC_CLUSTER / WB_DEC_VOCAB / WB_XATTN_TC = 1), so one process can hold contexts of either kind
    bool use_layers = false, use_vocab = false, use_xattn = false;
    bool vocab_ok = false;              // fused vocabulary kernel usable (whisper-base widths)
    bool xa_ok = false;                 // stand-alone tensor-core cross-attention usable (even key-tile count, head_dim 64)
    DevBuf<bf16> att, ffn;
    DevBuf<float> xpart;
    DevBuf<long long> prof;
    int cs = 0;                         // cluster size in use (16 or 8); 0 = path unavailable
    int max_clusters = 0;               // clusters of that size the GPU holds at once
};

void dec_cluster_free(wb_ctx* ctx) {
    delete ctx->dec.cluster;
    ctx->dec.cluster = nullptr;
}

// Called from decoder_alloc: builds the weight images, the layer table and the K/V tensor maps, picks the cluster size.
void dec_cluster_alloc(wb_ctx* ctx) {
    const wb_model_cfg& c = ctx->cfg;
    ctx->dec.cluster = nullptr;
    if (c.precision != WB_PREC_BF16) return;
    auto* dc = new DecCluster();
    ctx->dec.cluster = dc;
    auto env_on = [](const char* name) { const char* e = getenv(name); return e && e[0] == '1'; };
    dc->use_layers = env_on("WB_DEC_CLUSTER"); dc->use_vocab = env_on("WB_DEC_VOCAB"); dc->use_xattn = env_on("WB_XATTN_TC");

    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    CUDA_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    WB_REQUIRE(p != nullptr && q == cudaDriverEntryPointSuccess, WB_ECUDA, "cuTensorMapEncodeTiled not available");
    // K/V caches [layers][Bmax][T][2d] bf16 as 3-D tensors: tiles of 128 keys x one head (64 dims = 128 bytes, 128B swizzle);
    // keys past T are out of bounds in dimension 1 and arrive as zeros, never as the next sequence's rows
    auto make = [&](CUtensorMap* tm, void* base, int T) {
        cuuint64_t dims[3] = {(cuuint64_t)(2 * c.d_model), (cuuint64_t)T, (cuuint64_t)c.dec_layers * c.max_batch};
        cuuint64_t str[2] = {(cuuint64_t)(2 * c.d_model) * 2, (cuuint64_t)T * 2 * c.d_model * 2};
        cuuint32_t box[3] = {DC_HD, DC_KEYS, 1}, estr[3] = {1, 1, 1};
        CUresult r = reinterpret_cast<EncodeTiledFn>(p)(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, base, dims, str, box, estr,
                                                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        WB_REQUIRE(r == CUDA_SUCCESS, WB_ECUDA, "cuTensorMapEncodeTiled (decoder K/V cache) failed with CUresult %d", (int)r);
    };
    // ---- stand-alone tensor-core cross-attention: any bf16 model with head_dim 64 whose key tiles split into two halves ----
    const int nkt = ceil_div(c.n_audio_ctx, DC_KEYS);
    if (c.d_model == c.n_heads * DC_HD && nkt % 2 == 0 && nkt / 2 <= DC_MAXHALF) {
        make(&dc->tm_ckv, ctx->enc.ckv.p, c.n_audio_ctx);
        CUDA_CHECK(cudaFuncSetAttribute(cross_attn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, XA_SMEM));
        dc->xa_ok = true;
    }
    // ---- whisper-base widths only: fused vocabulary kernel and the (opt-in) cluster-chained layer kernel ----
    if (c.d_model != DC_D || c.n_heads != DC_H || c.ffn_dim != DC_FFN || !dc->xa_ok || c.n_text_ctx > 0xffff) return;
    dc->n_vchunks = ceil_div(c.vocab, DV_CROWS);
    CUDA_CHECK(cudaFuncSetAttribute(dec_vocab_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DV_SMEM));
    dc->pval.reserve((size_t)ctx->sm_count * DV_SEQ);
    dc->pidx.reserve((size_t)ctx->sm_count * DV_SEQ);
    dc->counter.reserve_zero(4);
    dc->vocab_ok = true;
    if (!dc->use_layers) return;                           // the layer kernel's weight images (47 MB) are built only when asked for
    int want = 16;
    if (const char* e = getenv("WB_DEC_CS")) want = atoi(e);
    const int n16 = want >= 16 ? max_active_clusters_of(dec_layers_kernel<16>, 16) : 0;
    const int n8 = max_active_clusters_of(dec_layers_kernel<8>, 8);
    if (n16 >= 1) { dc->cs = 16; dc->max_clusters = n16; }
    else if (n8 >= 1) { dc->cs = 8; dc->max_clusters = n8; }
    else return;
    if (getenv("WB_TRACE_CREATE")) {
        fprintf(stderr, "[dec_cluster] max active clusters by cluster size:");
        for (int cs = 1; cs <= 16; ++cs) fprintf(stderr, " %d:%d", cs, max_active_clusters_of(dc_occupancy_probe_kernel, cs));
        fprintf(stderr, "\n[dec_cluster] cluster size %d, %d clusters at once, %d B smem per CTA\n", dc->cs, dc->max_clusters, DC_SMEM);
    }
    std::vector<DcLayer> tab(c.dec_layers);
    dc->wimg.reserve((size_t)c.dec_layers * IMG_CHUNKS * DC_SLOT);
    CUDA_CHECK(cudaMemsetAsync(dc->wimg.p, 0, (size_t)c.dec_layers * IMG_CHUNKS * DC_SLOT, ctx->stream));
    for (int l = 0; l < c.dec_layers; ++l) {
        const DecLayerW& L = ctx->w.dec[l];
        DcLayer& t = tab[l];
        t.bqkv = L.qkv.b; t.bo = L.o.b; t.bcq = L.cq.b; t.bco = L.co.b; t.bfc1 = L.fc1.b; t.bfc2 = L.fc2.b;
        t.ln1w = L.ln1.w; t.ln1b = L.ln1.b; t.ln2w = L.ln2.w; t.ln2b = L.ln2.b; t.ln3w = L.ln3.w; t.ln3b = L.ln3.b;
        WB_REQUIRE(t.bqkv && t.bo && t.bcq && t.bco && t.bfc1 && t.bfc2, WB_EINVAL, "decoder layer %d: missing bias", l);
        bf16* img = reinterpret_cast<bf16*>(dc->wimg.p + (size_t)l * IMG_CHUNKS * DC_SLOT);
        auto pack = [&](const LinearW& W, int chunk0) {
            dc_pack_kernel<<<256, 256, 0, ctx->stream>>>((const bf16*)W.w, img + (size_t)chunk0 * DC_CROWS * DC_WPITCH, W.out, W.in);
        };
        pack(L.qkv, IMG_QKV); pack(L.o, IMG_O); pack(L.cq, IMG_CQ); pack(L.co, IMG_CO); pack(L.fc1, IMG_FC1); pack(L.fc2, IMG_FC2);
        CUDA_CHECK(cudaGetLastError());
    }
    dc->layers.reserve(sizeof(DcLayer) * tab.size());
    CUDA_CHECK(cudaMemcpyAsync(dc->layers.p, tab.data(), sizeof(DcLayer) * tab.size(), cudaMemcpyHostToDevice, ctx->stream));
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    if (const char* e = getenv("WB_DEC_PROF")) if (e[0] == '1') dc->prof.reserve_zero(512);
    dc->att.reserve((size_t)c.max_batch * DC_D);
    dc->ffn.reserve((size_t)c.max_batch * DC_FFN);
    dc->xpart.reserve((size_t)c.max_batch * DC_H * 2 * 66);
    make(&dc->tm_skv, ctx->dec.self_kv.p, ctx->dec.T_max);
    // the self-attention cache is read through the TMA unit: rows never written must not hold NaN bit patterns (a masked
    // key contributes 0 x value)
    CUDA_CHECK(cudaMemsetAsync(ctx->dec.self_kv.p, 0, (size_t)c.dec_layers * c.max_batch * ctx->dec.T_max * 2 * DC_D * 2, ctx->stream));
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
}

// The cluster-chained layer kernel is opt-in (WB_DEC_CLUSTER=1): measured on B200 it matches the per-kernel path on one
// batch (46 ms per 128-token decode of 32 sequences) but, holding 112 SMs for a whole step, it does not interleave with
// other batches in flight (21 k vs 33 k audio-s/s at 4 in flight); DESIGN.md has the stage profile.
bool dec_cluster_enabled(const wb_ctx* ctx) {
    return ctx->dec.cluster != nullptr && ctx->dec.cluster->use_layers && ctx->dec.cluster->cs > 0;
}

// All decoder layers of one step for sequences [0, B) (input token from prompt / cur_tok, output: x = residual stream
// after the last layer).  One launch.
void dec_cluster_layers(wb_ctx* ctx, cudaStream_t st, bool pdl, const int* state, const int* prompt_dev, const int* cur_tok, int B) {
    DecCluster* dc = ctx->dec.cluster;
    const wb_model_cfg& c = ctx->cfg;
    DcArgs a{};
    a.state = state; a.prompt = prompt_dev; a.cur_tok = cur_tok;
    a.E = (const bf16*)ctx->w.embed; a.P = ctx->w.dec_pos;
    a.layers = reinterpret_cast<const DcLayer*>(dc->layers.p); a.wimg = dc->wimg.p; a.n_layers = c.dec_layers;
    a.x = ctx->dec.x.p; a.qkv = ctx->dec.qkv.p; a.q = ctx->dec.q.p; a.att = dc->att.p; a.xpart = dc->xpart.p; a.ffn = dc->ffn.p;
    a.self_kv = reinterpret_cast<bf16*>(ctx->dec.self_kv.p);
    a.B = B; a.Bmax = c.max_batch; a.T_max = ctx->dec.T_max; a.Tk = c.n_audio_ctx;
    a.prof = dc->prof.p;            // null unless WB_DEC_PROF=1
    a.prof_stage = 5;
    if (const char* e = getenv("WB_DEC_PROF_STAGE")) a.prof_stage = atoi(e);
    // as many clusters as the GPU holds at once (one wave), at most one per sequence, at least enough for 8 sequences each
    int ncl = B < dc->max_clusters ? B : dc->max_clusters;
    if (const char* e = getenv("WB_DEC_NCL")) ncl = std::max(1, std::min(B, atoi(e)));
    if (ncl * DC_SEQ < B) ncl = ceil_div(B, DC_SEQ);
    a.n_clusters = ncl;
    cudaError_t e = dc->cs == 16 ? launch_layers<16>(st, pdl, ncl, dc->tm_ckv, dc->tm_skv, a)
                                 : launch_layers<8>(st, pdl, ncl, dc->tm_ckv, dc->tm_skv, a);
    CUDA_CHECK(e);
}

// Cross-attention of decoder layer l for sequences [0, B) on the tensor-map / ring kernel (bf16 build, head_dim 64).
bool cross_attn_tc_ok(const wb_ctx* ctx) {        // opt-in: 18.8 us vs 18.4 us for cross_attn_kernel<bf16> (both with PDL, B = 32)
    return ctx->dec.cluster != nullptr && ctx->dec.cluster->use_xattn && ctx->dec.cluster->xa_ok;
}
void cross_attn_tc(wb_ctx* ctx, cudaStream_t st, bool pdl, int layer, const float* q, float* out, int B) {
    DecCluster* dc = ctx->dec.cluster;
    const wb_model_cfg& c = ctx->cfg;
    XaArgs a{};
    a.q = q; a.out = out; a.H = c.n_heads; a.B = B; a.d = c.d_model; a.Tk = c.n_audio_ctx; a.zbase = layer * c.max_batch;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(std::min(ctx->sm_count, B * c.n_heads));
    cfg.blockDim = dim3(DC_THREADS);
    cfg.dynamicSmemBytes = XA_SMEM;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    CUDA_CHECK(cudaLaunchKernelEx(&cfg, cross_attn_tc_kernel, dc->tm_ckv, a));
}

// Final LayerNorm + vocabulary projection + arg-max + token bookkeeping + step advance for sequences [0, B), B <= 32.
bool dec_cluster_vocab_ok(const wb_ctx* ctx, int B) {       // opt-in: measured equal to the 3-launch path
    return ctx->dec.cluster != nullptr && ctx->dec.cluster->use_vocab && ctx->dec.cluster->vocab_ok && B <= DV_SEQ;
}
void dec_cluster_vocab(wb_ctx* ctx, cudaStream_t st, bool pdl, int* state, int B, float* logits, const int* forced, int max_new, int eot,
                       int T_total, int* cur_tok) {
    DecCluster* dc = ctx->dec.cluster;
    const wb_model_cfg& c = ctx->cfg;
    DecBufs& D = ctx->dec;
    DvArgs a{};
    a.state = state; a.x = D.x.p; a.lnw = ctx->w.dec_ln.w; a.lnb = ctx->w.dec_ln.b; a.E = (const bf16*)ctx->w.embed;
    a.V = c.vocab; a.B = B; a.n_chunks = dc->n_vchunks;
    a.sup_base = D.sup_base.p; a.sup_first = D.sup_first.p; a.logits = logits;
    a.pval = dc->pval.p; a.pidx = dc->pidx.p; a.counter = dc->counter.p;
    a.forced = forced; a.max_new = max_new; a.eot = eot; a.T_total = T_total;
    a.tokens = D.tokens.p; a.lens = D.lens.p; a.finished = D.finished.p; a.cur_tok = cur_tok; a.unfinished_out = D.unfinished_dev;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(std::min(ctx->sm_count, dc->n_vchunks));
    cfg.blockDim = dim3(DV_THREADS);
    cfg.dynamicSmemBytes = DV_SMEM;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    CUDA_CHECK(cudaLaunchKernelEx(&cfg, dec_vocab_kernel, a));
}
double dec_cluster_vocab_bytes(const wb_ctx* ctx) { return (double)ctx->cfg.vocab * DC_D * 2.0; }

// wb_bench_kernel("dec_layers"): algorithmic bytes of one launch, and (WB_DEC_PROF=1) the stage stamps of cluster 0.
double dec_cluster_bytes(const wb_ctx* ctx, int B) {
    const wb_model_cfg& c = ctx->cfg;
    const double w = (double)c.dec_layers * (14.0 * DC_D * DC_D) * 2.0;                    // qkv 3 + o + cq + co + fc1 4 + fc2 4 (x d^2), bf16
    const double kv = (double)c.dec_layers * B * 2.0 * c.n_audio_ctx * DC_D * 2.0;         // cached cross K/V of every sequence
    return w + kv;
}
void dec_cluster_print_prof(wb_ctx* ctx) {
    DecCluster* dc = ctx->dec.cluster;
    if (!dc || !dc->prof.p) return;
    const int n = 2 + 8 * ctx->cfg.dec_layers;
    std::vector<long long> h(n);
    CUDA_CHECK(cudaMemcpy(h.data(), dc->prof.p, sizeof(long long) * n, cudaMemcpyDeviceToHost));
    static const char* names[8] = {"ln1+qkv", "self_attn", "o_proj", "ln2+cq", "cross_attn", "co_proj", "ln3+fc1", "fc2"};
    double per[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int l = 0; l < ctx->cfg.dec_layers; ++l)
        for (int k = 0; k < 8; ++k) per[k] += (double)(h[2 + l * 8 + k] - h[1 + l * 8 + k]);
    fprintf(stderr, "[dec_cluster prof] cycles (cluster 0, CTA 0): prologue %lld, total %lld; per layer:", h[1] - h[0], h[n - 1] - h[0]);
    for (int k = 0; k < 8; ++k) fprintf(stderr, " %s %.0f", names[k], per[k] / ctx->cfg.dec_layers);
    fprintf(stderr, "\n");
    std::vector<long long> h2(200);
    CUDA_CHECK(cudaMemcpy(h2.data(), dc->prof.p + 256, sizeof(long long) * 200, cudaMemcpyDeviceToHost));
    fprintf(stderr, "[dec_cluster prof] fine stamps of layer 1 (stage WB_DEC_PROF_STAGE, default 5 = cross-attention), deltas:");
    for (int i = 1; i < 200 && h2[i] != 0; ++i) fprintf(stderr, " %lld", h2[i] - h2[i - 1]);
    fprintf(stderr, "\n");
    CUDA_CHECK(cudaMemset(dc->prof.p, 0, sizeof(long long) * 512));
}
