// encoder.cu — kernel group 2: the Whisper encoder forward (replaces run_encoder,
// /root/reference/src/main.rs:698-707, i.e. encoder.run at :703 on encoder_model.onnx) plus the
// cross-attention K/V projection that decoder_model.onnx emits as present.*.encoder.* (:786-787).
//
// Layout: activations are token-major [clip][frame][channel]; the conv stem is expressed as two
// GEMMs over overlapping row windows of a zero-padded time-major buffer (row t of conv1's A
// operand = frames t..t+2 of [3002][n_mels]; row t of conv2's = frames 2t..2t+2 of [3001][d]), so
// no im2col buffer ever touches HBM.  Residual stream f32; GEMM operands in the compute dtype.
#include "ctx.h"

namespace {

constexpr int LN_MAX_VEC = 10;        // float4 per lane: d_model <= 1280

__device__ __forceinline__ void put4(float* p, float a, float b, float c, float d) { *reinterpret_cast<float4*>(p) = make_float4(a, b, c, d); }
__device__ __forceinline__ void put4(__nv_bfloat16* p, float a, float b, float c, float d) {
    __nv_bfloat162 lo = __floats2bfloat162_rn(a, b), hi = __floats2bfloat162_rn(c, d);
    uint2 u;
    u.x = *reinterpret_cast<unsigned*>(&lo);
    u.y = *reinterpret_cast<unsigned*>(&hi);
    *reinterpret_cast<uint2*>(p) = u;
}

// One warp per row, 128-bit loads, row kept in registers (NV float4 per lane, d = 128*NV).
// out = (x-mean)/sqrt(var+eps)*w+b with two-pass statistics like the oracle; optional second f32 copy.
template <typename TO, int NV>
__global__ void __launch_bounds__(256)
layernorm_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
                 TO* __restrict__ out, float* __restrict__ out_f32, int rows) {
    constexpr int d = NV * 128;
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float4* xr = reinterpret_cast<const float4*>(x + (size_t)row * d);
    float4 v[NV];
    float s = 0.0f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        v[i] = xr[lane + i * 32];
        s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s / (float)d;
    float q = 0.0f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
        q += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rs = 1.0f / sqrtf(q / (float)d + 1e-5f);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int c = (lane + i * 32) * 4;
        const float4 g = __ldg(reinterpret_cast<const float4*>(w + c)), bb = __ldg(reinterpret_cast<const float4*>(b + c));
        const float y0 = v[i].x * rs * g.x + bb.x, y1 = v[i].y * rs * g.y + bb.y, y2 = v[i].z * rs * g.z + bb.z, y3 = v[i].w * rs * g.w + bb.w;
        put4(out + (size_t)row * d + c, y0, y1, y2, y3);
        if (out_f32) put4(out_f32 + (size_t)row * d + c, y0, y1, y2, y3);
    }
}

// Row softmax in place over f32 scores (SIMT attention path).  One warp per row, n <= 1536.
__global__ void softmax_rows_kernel(float* __restrict__ s, int rows, int n) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    float* r = s + (size_t)row * n;
    float v[48];
    float m = -INFINITY;
#pragma unroll
    for (int i = 0; i < 48; ++i) {
        int c = lane + i * 32;
        v[i] = c < n ? r[c] : -INFINITY;
        m = fmaxf(m, v[i]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float sum = 0.0f;
#pragma unroll
    for (int i = 0; i < 48; ++i) {
        int c = lane + i * 32;
        v[i] = c < n ? expf(v[i] - m) : 0.0f;
        sum += v[i];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
#pragma unroll
    for (int i = 0; i < 48; ++i) {
        int c = lane + i * 32;
        if (c < n) r[c] = v[i] / sum;
    }
}

template <typename TO>
void layernorm_t(wb_ctx* ctx, const float* x, const LNW& ln, TO* out, float* out_f32, int rows, int d) {
    const int wpb = 8;
    dim3 grid(ceil_div(rows, wpb));
    switch (d / 128) {
        case 1: layernorm_kernel<TO, 1><<<grid, wpb * 32, 0, ctx->stream>>>(x, ln.w, ln.b, out, out_f32, rows); break;
        case 2: layernorm_kernel<TO, 2><<<grid, wpb * 32, 0, ctx->stream>>>(x, ln.w, ln.b, out, out_f32, rows); break;
        case 3: layernorm_kernel<TO, 3><<<grid, wpb * 32, 0, ctx->stream>>>(x, ln.w, ln.b, out, out_f32, rows); break;
        case 4: layernorm_kernel<TO, 4><<<grid, wpb * 32, 0, ctx->stream>>>(x, ln.w, ln.b, out, out_f32, rows); break;
        case 6: layernorm_kernel<TO, 6><<<grid, wpb * 32, 0, ctx->stream>>>(x, ln.w, ln.b, out, out_f32, rows); break;
        case 8: layernorm_kernel<TO, 8><<<grid, wpb * 32, 0, ctx->stream>>>(x, ln.w, ln.b, out, out_f32, rows); break;
        case 10: layernorm_kernel<TO, 10><<<grid, wpb * 32, 0, ctx->stream>>>(x, ln.w, ln.b, out, out_f32, rows); break;
        default: WB_THROW(WB_EINVAL, "layernorm: unsupported d_model %d (128 x {1,2,3,4,6,8,10})", d);
    }
    CUDA_CHECK(cudaGetLastError());
}

void layernorm(wb_ctx* ctx, const float* x, const LNW& ln, void* out, float* out_f32, int rows) {
    const int d = ctx->cfg.d_model;
    if (ctx->cfg.precision == WB_PREC_BF16) layernorm_t<__nv_bfloat16>(ctx, x, ln, (__nv_bfloat16*)out, out_f32, rows, d);
    else layernorm_t<float>(ctx, x, ln, (float*)out, out_f32, rows, d);
}

}  // namespace

// SIMT attention (fp32 build, and bring-up reference for attn_tc.cu): S = QK^T, row softmax, PV
// through an f32 score scratch, in groups of attn_group clips.
int attention_simt(wb_ctx* ctx, const void* qkv_all, void* att_all, int B) {
    const wb_model_cfg& c = ctx->cfg;
    const int Tc = c.n_audio_ctx, d = c.d_model, H = c.n_heads;
    const int ct = c.precision == WB_PREC_BF16 ? WB_BF16 : WB_F32;
    EncBufs& b = ctx->enc;
    int launches = 0;
        for (int g0 = 0; g0 < B; g0 += b.attn_group) {
            const int G = (B - g0 < b.attn_group) ? B - g0 : b.attn_group;
            const char* qkv = (const char*)qkv_all + (size_t)g0 * Tc * 3 * d * ctx->esz();
            {   // S = (q * hd^-0.5) k^T   (0.125 is a power of two: exact in either order)
                GemmArgs g;
                g.A = qkv; g.B = qkv + (size_t)d * ctx->esz(); g.C = b.scores.p;
                g.ta = g.tb = ct; g.tc = WB_F32;
                g.M = Tc; g.N = Tc; g.K = 64; g.lda = 3 * d; g.ldb = 3 * d; g.ldc = Tc;
                g.batch = G * H; g.inner = H;
                g.sAo = (long long)Tc * 3 * d; g.sAi = 64; g.sBo = g.sAo; g.sBi = 64;
                g.sCo = (long long)H * Tc * Tc; g.sCi = (long long)Tc * Tc;
                g.alpha = 0.125f;
                gemm_simt(ctx, g); ++launches;
            }
            softmax_rows_kernel<<<ceil_div(G * H * Tc, 8), 256, 0, ctx->stream>>>(b.scores.p, G * H * Tc, Tc);
            CUDA_CHECK(cudaGetLastError()); ++launches;
            {   // O = P v
                GemmArgs g;
                g.A = b.scores.p; g.B = qkv + (size_t)2 * d * ctx->esz();
                g.C = (char*)att_all + (size_t)g0 * Tc * d * ctx->esz();
                g.ta = WB_F32; g.tb = ct; g.tc = ct; g.b_kn = true;
                g.M = Tc; g.N = 64; g.K = Tc; g.lda = Tc; g.ldb = 3 * d; g.ldc = d;
                g.batch = G * H; g.inner = H;
                g.sAo = (long long)H * Tc * Tc; g.sAi = (long long)Tc * Tc;
                g.sBo = (long long)Tc * 3 * d; g.sBi = 64;
                g.sCo = (long long)Tc * d; g.sCi = 64;
                gemm_simt(ctx, g); ++launches;
            }
        }
    return launches;
}

void encoder_alloc(wb_ctx* ctx) {
    const wb_model_cfg& c = ctx->cfg;
    const size_t B = c.max_batch, T = WB_N_FRAMES, Tc = c.n_audio_ctx, d = c.d_model, e = ctx->esz();
    WB_REQUIRE(c.n_audio_ctx * 2 == WB_N_FRAMES, WB_EINVAL, "n_audio_ctx must be 1500");
    WB_REQUIRE(d <= 128 * LN_MAX_VEC && d % 128 == 0 && c.ffn_dim % 128 == 0, WB_EINVAL, "unsupported d_model/ffn_dim");
    WB_REQUIRE(d / c.n_heads == 64 && d % c.n_heads == 0, WB_EINVAL, "head_dim must be 64");
    EncBufs& b = ctx->enc;
    b.mel_tm.reserve_zero((size_t)c.max_chunks * (T + 2) * c.n_mels * e);
    b.in_tm.reserve_zero(B * (T + 2) * c.n_mels * e);
    b.in_stage.reserve(B * c.n_mels * T);
    b.h1p.reserve_zero(B * (T + 1) * d * e);
    b.x.reserve(B * Tc * d);
    b.h.reserve(B * Tc * d * e);
    b.qkv.reserve(B * Tc * 3 * d * e);
    b.att.reserve(B * Tc * d * e);
    b.ffn.reserve(B * Tc * c.ffn_dim * e);
    b.attn_group = (int)(B < 4 ? B : 4);
    b.scores.reserve((size_t)b.attn_group * c.n_heads * Tc * Tc);
    b.out.reserve(B * Tc * d);
    if (c.precision == WB_PREC_BF16) b.out_c.reserve(B * Tc * d * e);
    b.ckv.reserve((size_t)c.dec_layers * B * Tc * 2 * d * e);
}

void encoder_forward(wb_ctx* ctx, const void* mel_tm, int B) {
    const wb_model_cfg& c = ctx->cfg;
    const int T = WB_N_FRAMES, Tc = c.n_audio_ctx, d = c.d_model, H = c.n_heads, C = c.n_mels;
    const int ct = c.precision == WB_PREC_BF16 ? WB_BF16 : WB_F32;
    EncBufs& b = ctx->enc;
    ModelW& w = ctx->w;
    CUDA_CHECK(cudaEventRecord(ctx->enc_e0.e, ctx->stream));
    int launches = 0;

    {   // conv1 + GELU -> h1p rows 1..3000           (K2a)
        GemmArgs g;
        g.A = mel_tm; g.B = w.conv1.w; g.C = (char*)b.h1p.p + (size_t)d * ctx->esz();
        g.ta = g.tb = g.tc = ct;
        g.M = T; g.N = d; g.K = 3 * C; g.lda = C; g.ldb = 3 * C; g.ldc = d;
        g.batch = B; g.inner = 1; g.sAo = (long long)(T + 2) * C; g.sCo = (long long)(T + 1) * d;
        g.bias = w.conv1.b; g.act = 1;
        gemm(ctx, g); ++launches;
    }
    {   // conv2 (stride 2) + GELU + positions -> x
        GemmArgs g;
        g.A = b.h1p.p; g.B = w.conv2.w; g.C = b.x.p;
        g.ta = g.tb = ct; g.tc = WB_F32;
        g.M = Tc; g.N = d; g.K = 3 * d; g.lda = 2 * d; g.ldb = 3 * d; g.ldc = d;
        g.batch = B; g.inner = 1; g.sAo = (long long)(T + 1) * d; g.sCo = (long long)Tc * d;
        g.bias = w.conv2.b; g.act = 1; g.rowadd = w.enc_pos; g.ld_rowadd = d;
        gemm(ctx, g); ++launches;
    }
    if (ctx->debug) {
        b.dbg_stem.reserve((size_t)c.max_batch * Tc * d);
        CUDA_CHECK(cudaMemcpyAsync(b.dbg_stem.p, b.x.p, sizeof(float) * (size_t)B * Tc * d, cudaMemcpyDeviceToDevice, ctx->stream));
    }

    const int rows = B * Tc;
    for (int l = 0; l < c.enc_layers; ++l) {
        const EncLayerW& L = w.enc[l];
        layernorm(ctx, b.x.p, L.ln1, b.h.p, nullptr, rows); ++launches;                  // K2b
        {   // fused QKV projection, N = 3d                                             (K2c)
            GemmArgs g;
            g.A = b.h.p; g.B = L.qkv.w; g.C = b.qkv.p; g.ta = g.tb = g.tc = ct;
            g.M = rows; g.N = 3 * d; g.K = d; g.lda = d; g.ldb = d; g.ldc = 3 * d; g.bias = L.qkv.b;
            gemm(ctx, g); ++launches;
        }
        // self-attention, non-causal, T = 1500, head_dim 64                            (K2d)
        if (ct == WB_BF16 && attn_tc_enabled()) { attn_tc(ctx, b.qkv.p, b.att.p, B, Tc, d, H); ++launches; }
        else launches += attention_simt(ctx, b.qkv.p, b.att.p, B);
        {   // out-proj + bias + residual (in place on x)
            GemmArgs g;
            g.A = b.att.p; g.B = L.o.w; g.C = b.x.p; g.ta = g.tb = ct; g.tc = WB_F32;
            g.M = rows; g.N = d; g.K = d; g.lda = d; g.ldb = d; g.ldc = d; g.bias = L.o.b; g.residual = b.x.p;
            gemm(ctx, g); ++launches;
        }
        layernorm(ctx, b.x.p, L.ln2, b.h.p, nullptr, rows); ++launches;
        {   // fc1 + GELU
            GemmArgs g;
            g.A = b.h.p; g.B = L.fc1.w; g.C = b.ffn.p; g.ta = g.tb = g.tc = ct;
            g.M = rows; g.N = c.ffn_dim; g.K = d; g.lda = d; g.ldb = d; g.ldc = c.ffn_dim; g.bias = L.fc1.b; g.act = 1;
            gemm(ctx, g); ++launches;
        }
        {   // fc2 + bias + residual
            GemmArgs g;
            g.A = b.ffn.p; g.B = L.fc2.w; g.C = b.x.p; g.ta = g.tb = ct; g.tc = WB_F32;
            g.M = rows; g.N = d; g.K = c.ffn_dim; g.lda = c.ffn_dim; g.ldb = c.ffn_dim; g.ldc = d; g.bias = L.fc2.b; g.residual = b.x.p;
            gemm(ctx, g); ++launches;
        }
        if (ctx->debug && l == 0) {
            b.dbg_layer0.reserve((size_t)c.max_batch * Tc * d);
            CUDA_CHECK(cudaMemcpyAsync(b.dbg_layer0.p, b.x.p, sizeof(float) * (size_t)B * Tc * d, cudaMemcpyDeviceToDevice, ctx->stream));
        }
    }
    // final LayerNorm -> f32 output (ONNX output 0) + compute-dtype copy for the cross-K/V GEMMs
    const void* enc_c;
    if (ct == WB_BF16) {
        layernorm(ctx, b.x.p, w.enc_ln, b.out_c.p, b.out.p, rows);
        enc_c = b.out_c.p;
    } else {
        layernorm(ctx, b.x.p, w.enc_ln, b.out.p, nullptr, rows);
        enc_c = b.out.p;
    }
    ++launches;
    CUDA_CHECK(cudaEventRecord(ctx->enc_e1.e, ctx->stream));

    // cross-attention K/V for every decoder layer: [layer][B][1500][2d] (k | v)       (K3a)
    for (int l = 0; l < c.dec_layers; ++l) {
        GemmArgs g;
        g.A = enc_c; g.B = w.dec[l].ckv.w;
        g.C = (char*)b.ckv.p + (size_t)l * c.max_batch * Tc * 2 * d * ctx->esz();
        g.ta = g.tb = g.tc = ct;
        g.M = rows; g.N = 2 * d; g.K = d; g.lda = d; g.ldb = d; g.ldc = 2 * d; g.bias = w.dec[l].ckv.b;
        gemm(ctx, g); ++launches;
    }
    CUDA_CHECK(cudaEventRecord(ctx->enc_e2.e, ctx->stream));
    ctx->enc_t_pending = true;            // encoder_ms / cross_kv_ms are read back by timing_flush: the decode is enqueued behind
    ctx->timing.encoder_launches = launches;   // this without a host wait
    b.B_valid = B;
}

void gemm(wb_ctx* ctx, const GemmArgs& a) {
    if (gemm_tc_eligible(a)) gemm_tc(ctx, a);
    else gemm_simt(ctx, a);
}
