// field_mlp.cu -- the density + endmember-abundance MLP heads of UMHSField, fused (fp32 SIMT path).
//
// Replaces, per sample (ref umhsnerf/umhs_field.py):
//   :320-328  base MLP 32->64->16, trunc_exp, selector            -> sigma, emb[15]
//   :183-184  NeRF position encoding of the WORLD position        -> pe[12]
//   :211-212  mlp_head 27->64->64->K, sigmoid                     -> scalar[K]
//   :214-226  feature_mlp 27->64->64->K(+1), softmax(/T), sigmoid -> abund[K], s1
//   :242-250  mlp_directional first layer [SH16,pe12]->16, ReLU   -> spec_hidden[16]
// and the autograd backward of all of it (dgrad to the hash encoding + wgrad of every layer).
//
// Layout: one thread per sample.  Layer weights are staged in shared memory ([out][in_pad] fp32,
// read as warp-broadcast float4); each thread keeps its sample's activations in a private
// shared-memory row (stride 68/36 words: float4 accesses are bank-conflict free) so that the
// weight gradients can be formed CTA-wide as dW[j][i] = sum_s D[s][j] * A[s][i] from the same
// rows, register-blocked 1x32 per thread, one atomic per element per 128-sample tile.
// Activations are recomputed in the backward pass from the hash encoding (nothing but
// enc[N,32] and emb[N,15] is kept from the forward).
#include "common.cuh"

namespace umhs {

constexpr int H = 64;        // hidden width
constexpr int ENC = 32;      // hash encoding width (16 levels x 2)
constexpr int XW = 28;       // 27 (pe12 + emb15) or 28 (sh16 + pe12), padded to a float4 multiple
constexpr int LDH = 68;      // row stride of a 64-wide activation row
constexpr int LDX = 36;      // row stride of a 32-wide row
constexpr int OUTP = 24;     // padded width of the last-layer rows (K+1 <= 17 -> 24)

struct WOff {  // offsets (floats) of one 3-layer MLP in smem
    int w0, b0, w1, b1, w2, b2, total;
};
__host__ __device__ inline WOff mlp3_offsets(int nout) {
    WOff o;
    int p = 0;
    o.w0 = p; p += H * XW;
    o.b0 = p; p += H;
    o.w1 = p; p += H * H;
    o.b1 = p; p += H;
    o.w2 = p; p += ((nout + 3) / 4 * 4) * H;
    o.b2 = p; p += OUTP;
    o.total = p;
    return o;
}
constexpr int kBaseW0 = 0, kBaseB0 = H * ENC, kBaseW1 = kBaseB0 + H, kBaseB1 = kBaseW1 + 16 * H,
              kBaseTotal = kBaseB1 + 16;
constexpr int kDirW0 = 0, kDirB0 = 16 * XW, kDirTotal = kDirB0 + 16;

// cooperative copy of an nn.Linear weight [out][in] into smem [out_pad][in_pad] (zero padded)
__device__ __forceinline__ void stage_w(float *dst, const float *src, int out, int in, int out_pad, int in_pad,
                                        int tid, int nthr) {
    for (int i = tid; i < out_pad * in_pad; i += nthr) {
        const int j = i / in_pad, k = i - j * in_pad;
        dst[i] = (j < out && k < in) ? __ldg(src + j * in + k) : 0.f;
    }
}
__device__ __forceinline__ void stage_b(float *dst, const float *src, int n, int n_pad, int tid, int nthr) {
    for (int i = tid; i < n_pad; i += nthr) dst[i] = i < n ? __ldg(src + i) : 0.f;
}

template <int IN>
__device__ __forceinline__ void load_row(const float *row, float (&v)[IN]) {
#pragma unroll
    for (int i = 0; i < IN; i += 4) {
        const float4 t = *reinterpret_cast<const float4 *>(row + i);
        v[i] = t.x; v[i + 1] = t.y; v[i + 2] = t.z; v[i + 3] = t.w;
    }
}

// out_row[j] = act(b[j] + sum_i W[j*IN+i] * in[i]) for j < nout (nout multiple of 4, runtime)
template <int IN, bool RELU>
__device__ __forceinline__ void dense_row(const float *__restrict__ W, const float *__restrict__ b,
                                          const float (&in)[IN], float *out_row, int nout) {
    for (int j = 0; j < nout; j += 4) {
        float a0 = b[j], a1 = b[j + 1], a2 = b[j + 2], a3 = b[j + 3];
        const float *w = W + j * IN;
#pragma unroll
        for (int i = 0; i < IN; i += 4) {
            const float4 w0 = *reinterpret_cast<const float4 *>(w + i);
            const float4 w1 = *reinterpret_cast<const float4 *>(w + IN + i);
            const float4 w2 = *reinterpret_cast<const float4 *>(w + 2 * IN + i);
            const float4 w3 = *reinterpret_cast<const float4 *>(w + 3 * IN + i);
            a0 += w0.x * in[i] + w0.y * in[i + 1] + w0.z * in[i + 2] + w0.w * in[i + 3];
            a1 += w1.x * in[i] + w1.y * in[i + 1] + w1.z * in[i + 2] + w1.w * in[i + 3];
            a2 += w2.x * in[i] + w2.y * in[i + 1] + w2.z * in[i + 2] + w2.w * in[i + 3];
            a3 += w3.x * in[i] + w3.y * in[i + 1] + w3.z * in[i + 2] + w3.w * in[i + 3];
        }
        if (RELU) {
            a0 = fmaxf(a0, 0.f); a1 = fmaxf(a1, 0.f); a2 = fmaxf(a2, 0.f); a3 = fmaxf(a3, 0.f);
        }
        *reinterpret_cast<float4 *>(out_row + j) = make_float4(a0, a1, a2, a3);
    }
}

// acc[i] = sum_{j<nout} W[j*IN+i] * dz_row[j]   (dgrad through one Linear, thread-private)
template <int IN>
__device__ __forceinline__ void dgrad_row(const float *__restrict__ W, const float *dz_row, int nout,
                                          float (&acc)[IN]) {
#pragma unroll
    for (int i = 0; i < IN; ++i) acc[i] = 0.f;
    for (int j = 0; j < nout; ++j) {
        const float d = dz_row[j];
        const float *w = W + j * IN;
#pragma unroll
        for (int i = 0; i < IN; i += 4) {
            const float4 t = *reinterpret_cast<const float4 *>(w + i);
            acc[i] += t.x * d; acc[i + 1] += t.y * d; acc[i + 2] += t.z * d; acc[i + 3] += t.w * d;
        }
    }
}

// CTA-wide (128 threads) weight gradient: G[p][q] += sum_{s<nsamp} P[s*ldP+p] * Q[s*ldQ+q],
// p < nP (<= 64, one per thread), q < nQ (<= 2*NI); element (p,q) lives at dW[p*sP + q*sQ].
// dbP (optional): dbP[p] += sum_s P[s][p].
template <int NI>
__device__ __forceinline__ void wgrad_tile(const float *P, int ldP, int nP, const float *Q, int ldQ, int nQ,
                                           int nsamp, float *dW, int sP, int sQ, float *dbP) {
    const int p = threadIdx.x & 63, qh = threadIdx.x >> 6, q0 = qh * NI;
    float acc[NI], accb = 0.f;
#pragma unroll
    for (int q = 0; q < NI; ++q) acc[q] = 0.f;
    for (int s = 0; s < nsamp; ++s) {
        const float a = P[s * ldP + p];
        accb += a;
        const float *qr = Q + s * ldQ + q0;
#pragma unroll
        for (int q = 0; q < NI; q += 4) {
            const float4 t = *reinterpret_cast<const float4 *>(qr + q);
            acc[q] += a * t.x; acc[q + 1] += a * t.y; acc[q + 2] += a * t.z; acc[q + 3] += a * t.w;
        }
    }
    if (p < nP) {
#pragma unroll
        for (int q = 0; q < NI; ++q)
            if (q0 + q < nQ && acc[q] != 0.f) atomicAdd(dW + p * sP + (q0 + q) * sQ, acc[q]);
        if (dbP && qh == 0 && accb != 0.f) atomicAdd(dbP + p, accb);
    }
}

// column sums of a narrow row block: db[j] += sum_s D[s*ld + j], j < n (n <= 32)
__device__ __forceinline__ void colsum_tile(const float *D, int ld, int n, int nsamp, float *db) {
    if ((int)threadIdx.x < n) {
        float a = 0.f;
        for (int s = 0; s < nsamp; ++s) a += D[s * ld + threadIdx.x];
        if (a != 0.f) atomicAdd(db + threadIdx.x, a);
    }
}

__device__ __forceinline__ void pos_encoding(const Point &p, float (&pe)[12]) {
    // NeRFEncoding(in_dim=3, num_frequencies=2, min_freq_exp=0, max_freq_exp=1), include_input=False
    const float two_pi = 6.283185307179586f, half_pi = 1.5707963267948966f;
    const float s[3] = {two_pi * p.wx, two_pi * p.wy, two_pi * p.wz};
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const float a = s[d], b = s[d] * 2.0f;
        pe[2 * d] = sinf(a);
        pe[2 * d + 1] = sinf(b);
        pe[6 + 2 * d] = sinf(a + half_pi);
        pe[6 + 2 * d + 1] = sinf(b + half_pi);
    }
}

__device__ __forceinline__ void sh_deg3(float x, float y, float z, float (&c)[16]) {
    const float xx = x * x, yy = y * y, zz = z * z;
    c[0] = 0.28209479177387814f;
    c[1] = 0.4886025119029199f * y;
    c[2] = 0.4886025119029199f * z;
    c[3] = 0.4886025119029199f * x;
    c[4] = 1.0925484305920792f * x * y;
    c[5] = 1.0925484305920792f * y * z;
    c[6] = 0.9461746957575601f * zz - 0.31539156525251999f;
    c[7] = 1.0925484305920792f * x * z;
    c[8] = 0.5462742152960396f * (xx - yy);
    c[9] = 0.5900435899266435f * y * (3.0f * xx - yy);
    c[10] = 2.890611442640554f * x * y * z;
    c[11] = 0.4570457994644658f * y * (5.0f * zz - 1.0f);
    c[12] = 0.3731763325901154f * z * (5.0f * zz - 3.0f);
    c[13] = 0.4570457994644658f * x * (5.0f * zz - 1.0f);
    c[14] = 1.445305721320277f * z * (xx - yy);
    c[15] = 0.5900435899266435f * x * (xx - 3.0f * yy);
}

// ------------------------------------------------------------------------------- forward
constexpr int kFwdThreads = 256;

struct FwdSmem {
    int base, head, feat, dir, bufA, bufB, total;
};
__host__ __device__ inline FwdSmem fwd_layout(int K, int specular, bool density_only, int nthr) {
    FwdSmem s;
    int p = 0;
    s.base = p; p += kBaseTotal;
    s.head = p; p += density_only ? 0 : mlp3_offsets(K).total;
    s.feat = p; p += density_only ? 0 : mlp3_offsets(K + (specular ? 1 : 0)).total;
    s.dir = p; p += (!density_only && specular) ? kDirTotal : 0;
    s.bufA = p; p += nthr * LDH;
    s.bufB = p; p += density_only ? 0 : nthr * LDH;
    s.total = p;
    return s;
}

__device__ __forceinline__ void stage_mlp3(float *dst, const WOff &o, const float *w0, const float *b0,
                                           const float *w1, const float *b1, const float *w2, const float *b2, int nout,
                                           int tid, int nthr) {
    stage_w(dst + o.w0, w0, H, 27, H, XW, tid, nthr);
    stage_b(dst + o.b0, b0, H, H, tid, nthr);
    stage_w(dst + o.w1, w1, H, H, H, H, tid, nthr);
    stage_b(dst + o.b1, b1, H, H, tid, nthr);
    stage_w(dst + o.w2, w2, nout, H, (nout + 3) / 4 * 4, H, tid, nthr);
    stage_b(dst + o.b2, b2, nout, OUTP, tid, nthr);
}

template <bool DENSITY_ONLY>
__global__ void __launch_bounds__(kFwdThreads)
field_mlp_fwd_kernel(umhs_points P, int64_t n_total, const float *__restrict__ enc, umhs_mlp_params mp,
                     umhs_field_out out) {
    extern __shared__ __align__(16) float smem[];
    const int K = mp.num_classes, spec = mp.specular;
    const int tid = threadIdx.x;
    const FwdSmem L = fwd_layout(K, spec, DENSITY_ONLY, kFwdThreads);
    const WOff oh = mlp3_offsets(K), of = mlp3_offsets(K + (spec ? 1 : 0));
    float *Wb = smem + L.base, *Wh = smem + L.head, *Wf = smem + L.feat, *Wd = smem + L.dir;
    stage_w(Wb + kBaseW0, mp.base_w0, H, ENC, H, ENC, tid, kFwdThreads);
    stage_b(Wb + kBaseB0, mp.base_b0, H, H, tid, kFwdThreads);
    stage_w(Wb + kBaseW1, mp.base_w1, 16, H, 16, H, tid, kFwdThreads);
    stage_b(Wb + kBaseB1, mp.base_b1, 16, 16, tid, kFwdThreads);
    if (!DENSITY_ONLY) {
        stage_mlp3(Wh, oh, mp.head_w0, mp.head_b0, mp.head_w1, mp.head_b1, mp.head_w2, mp.head_b2, K, tid, kFwdThreads);
        stage_mlp3(Wf, of, mp.feat_w0, mp.feat_b0, mp.feat_w1, mp.feat_b1, mp.feat_w2, mp.feat_b2, K + (spec ? 1 : 0),
                   tid, kFwdThreads);
        if (spec) {
            stage_w(Wd + kDirW0, mp.dir_w0, 16, 28, 16, XW, tid, kFwdThreads);
            stage_b(Wd + kDirB0, mp.dir_b0, 16, 16, tid, kFwdThreads);
        }
    }
    __syncthreads();
    float *rowA = smem + L.bufA + tid * LDH;
    float *rowB = smem + L.bufB + tid * LDH;
    const float invT = 1.0f / mp.temperature;

    n_total = effective_n(P, n_total);
    for (int64_t n = (int64_t)blockIdx.x * kFwdThreads + tid; n < n_total; n += (int64_t)gridDim.x * kFwdThreads) {
        const Point pt = eval_point(P, n);
        float o16[16];
        {
            float in[ENC];
            const float4 *e4 = reinterpret_cast<const float4 *>(enc + enc_row(P, n) * ENC);
#pragma unroll
            for (int i = 0; i < ENC / 4; ++i) {
                const float4 t = __ldg(e4 + i);
                in[4 * i] = t.x; in[4 * i + 1] = t.y; in[4 * i + 2] = t.z; in[4 * i + 3] = t.w;
            }
            dense_row<ENC, true>(Wb + kBaseW0, Wb + kBaseB0, in, rowA, H);
        }
        {
            float h[H];
            load_row<H>(rowA, h);
            dense_row<H, false>(Wb + kBaseW1, Wb + kBaseB1, h, rowA, 16);
            load_row<16>(rowA, o16);
        }
        // trunc_exp forward is a plain exp; density * selector (ref umhs_field.py:327-328)
        out.sigma[n] = expf(o16[0]) * pt.sel;
        if (out.embedding) {
#pragma unroll
            for (int i = 0; i < 15; ++i) out.embedding[n * 15 + i] = o16[i + 1];
        }
        if (DENSITY_ONLY) continue;

        float x[XW], pe[12];
        pos_encoding(pt, pe);
#pragma unroll
        for (int i = 0; i < 12; ++i) x[i] = pe[i];
#pragma unroll
        for (int i = 0; i < 15; ++i) x[12 + i] = o16[1 + i];
        x[27] = 0.f;
        dense_row<XW, true>(Wh + oh.w0, Wh + oh.b0, x, rowA, H);
        dense_row<XW, true>(Wf + of.w0, Wf + of.b0, x, rowB, H);
        {
            float h[H];
            load_row<H>(rowA, h);
            dense_row<H, true>(Wh + oh.w1, Wh + oh.b1, h, rowA, H);
            load_row<H>(rowA, h);
            dense_row<H, false>(Wh + oh.w2, Wh + oh.b2, h, rowA, (K + 3) / 4 * 4);
            for (int k = 0; k < K; ++k) out.scalar[n * K + k] = 1.0f / (1.0f + expf(-rowA[k]));
            load_row<H>(rowB, h);
            dense_row<H, true>(Wf + of.w1, Wf + of.b1, h, rowB, H);
            load_row<H>(rowB, h);
            dense_row<H, false>(Wf + of.w2, Wf + of.b2, h, rowB, (K + spec + 3) / 4 * 4);
        }
        {
            float mx = -3.0e38f;
            for (int k = 0; k < K; ++k) mx = fmaxf(mx, rowB[k] * invT);
            float den = 0.f;
            for (int k = 0; k < K; ++k) {
                const float e = expf(rowB[k] * invT - mx);
                rowA[k] = e;
                den += e;
            }
            for (int k = 0; k < K; ++k) out.abund[n * K + k] = rowA[k] / den;
            if (spec && out.s1) out.s1[n] = 1.0f / (1.0f + expf(-rowB[K]));
        }
        if (spec && out.spec_hidden) {
            float sh[16];
            sh_deg3((pt.dx + 1.0f) / 2.0f, (pt.dy + 1.0f) / 2.0f, (pt.dz + 1.0f) / 2.0f, sh);
#pragma unroll
            for (int i = 0; i < 16; ++i) x[i] = sh[i];
#pragma unroll
            for (int i = 0; i < 12; ++i) x[16 + i] = pe[i];
            dense_row<XW, true>(Wd + kDirW0, Wd + kDirB0, x, rowA, 16);
            float4 *dst = reinterpret_cast<float4 *>(out.spec_hidden + n * 16);
#pragma unroll
            for (int i = 0; i < 4; ++i) dst[i] = *reinterpret_cast<const float4 *>(rowA + 4 * i);
        }
    }
}

// ------------------------------------------------------------------------------- backward
constexpr int kBwdThreads = 128;

struct BwdSmem {
    int w, X, H1, H2, D1, D2, total;
};
__host__ __device__ inline BwdSmem bwd_layout(int K, int specular) {
    BwdSmem s;
    int p = 0;
    int wmax = mlp3_offsets(K + (specular ? 1 : 0)).total;
    if (wmax < kBaseTotal) wmax = kBaseTotal;
    s.w = p; p += wmax;
    s.X = p; p += kBwdThreads * LDX;
    s.H1 = p; p += kBwdThreads * LDH;
    s.H2 = p; p += kBwdThreads * LDH;
    s.D1 = p; p += kBwdThreads * LDH;
    s.D2 = p; p += kBwdThreads * LDH;
    s.total = p;
    return s;
}

// One 27->64->64->nout MLP: recompute forward from x, form dz via `OutGrad`, wgrad of all three layers
// (atomics into g*), and add the gradient w.r.t. x[12..26] (the embedding) into demb[15].
// mode 0: scalar = sigmoid(z);  mode 1: abund = softmax(z[:K]/T), s1 = sigmoid(z[K]).
__device__ __forceinline__ void mlp3_backward(int mode, int K, int spec, float invT, bool valid, int nsamp, int64_t n,
                                              const float (&x)[XW], float *W, const WOff &o, float *rowX, float *rowH1,
                                              float *rowH2, float *rowD1, float *rowD2, const float *Xs,
                                              const float *H1s, const float *H2s, const float *D1s, const float *D2s,
                                              const float *d_out, const float *d_s1, float *gw0, float *gb0,
                                              float *gw1, float *gb1, float *gw2, float *gb2, float (&demb)[15]) {
    const int nout = (mode == 1 && spec) ? K + 1 : K;
    const int noutp = (nout + 3) / 4 * 4;
    // ---- forward recompute (thread-private rows)
#pragma unroll
    for (int i = 0; i < XW; i += 4) *reinterpret_cast<float4 *>(rowX + i) = make_float4(x[i], x[i + 1], x[i + 2], x[i + 3]);
    dense_row<XW, true>(W + o.w0, W + o.b0, x, rowH1, H);
    {
        float h[H];
        load_row<H>(rowH1, h);
        dense_row<H, true>(W + o.w1, W + o.b1, h, rowH2, H);
        load_row<H>(rowH2, h);
        dense_row<H, false>(W + o.w2, W + o.b2, h, rowD2, noutp);
    }
    // ---- dz into rowD2[0..OUTP)
    if (valid) {
        if (mode == 0) {
            for (int k = 0; k < K; ++k) {
                const float s = 1.0f / (1.0f + expf(-rowD2[k]));
                rowD2[k] = __ldg(d_out + n * K + k) * s * (1.0f - s);
            }
        } else {
            float mx = -3.0e38f;
            for (int k = 0; k < K; ++k) mx = fmaxf(mx, rowD2[k] * invT);
            float den = 0.f, dot = 0.f;
            for (int k = 0; k < K; ++k) den += expf(rowD2[k] * invT - mx);
            for (int k = 0; k < K; ++k) dot += (expf(rowD2[k] * invT - mx) / den) * __ldg(d_out + n * K + k);
            const float zs = spec ? rowD2[K] : 0.f;
            for (int k = 0; k < K; ++k) {
                const float a = expf(rowD2[k] * invT - mx) / den;
                rowD2[k] = a * invT * (__ldg(d_out + n * K + k) - dot);
            }
            if (spec) {
                const float s = 1.0f / (1.0f + expf(-zs));
                rowD2[K] = (d_s1 ? __ldg(d_s1 + n) : 0.f) * s * (1.0f - s);
            }
        }
        for (int k = nout; k < OUTP; ++k) rowD2[k] = 0.f;
    } else {
        for (int k = 0; k < OUTP; ++k) rowD2[k] = 0.f;
    }
    __syncthreads();
    // ---- layer 2: dW2[j][i] = sum_s D2[s][j] H2[s][i]  (transposed roles: p = i, q = j)
    wgrad_tile<12>(H2s, LDH, H, D2s, LDH, nout, nsamp, gw2, 1, H, nullptr);
    colsum_tile(D2s, LDH, nout, nsamp, gb2);
    {
        float acc[H];
        dgrad_row<H>(W + o.w2, rowD2, nout, acc);
#pragma unroll
        for (int i = 0; i < H; i += 4) {
            const float4 h = *reinterpret_cast<const float4 *>(rowH2 + i);
            *reinterpret_cast<float4 *>(rowD1 + i) = make_float4(h.x > 0.f ? acc[i] : 0.f, h.y > 0.f ? acc[i + 1] : 0.f,
                                                                 h.z > 0.f ? acc[i + 2] : 0.f, h.w > 0.f ? acc[i + 3] : 0.f);
        }
    }
    __syncthreads();
    // ---- layer 1: dW1[j][i] = sum_s D1[s][j] H1[s][i]
    wgrad_tile<32>(D1s, LDH, H, H1s, LDH, H, nsamp, gw1, H, 1, gb1);
    __syncthreads();  // D2 rows are rewritten below while other warps may still read them in colsum/wgrad above
    {
        float acc[H];
        dgrad_row<H>(W + o.w1, rowD1, H, acc);
#pragma unroll
        for (int i = 0; i < H; i += 4) {
            const float4 h = *reinterpret_cast<const float4 *>(rowH1 + i);
            *reinterpret_cast<float4 *>(rowD2 + i) = make_float4(h.x > 0.f ? acc[i] : 0.f, h.y > 0.f ? acc[i + 1] : 0.f,
                                                                 h.z > 0.f ? acc[i + 2] : 0.f, h.w > 0.f ? acc[i + 3] : 0.f);
        }
    }
    __syncthreads();
    // ---- layer 0: dW0[j][i] = sum_s D2[s][j] X[s][i], i < 27
    wgrad_tile<16>(D2s, LDH, H, Xs, LDX, 27, nsamp, gw0, 27, 1, gb0);
    {
        float acc[XW];
        dgrad_row<XW>(W + o.w0, rowD2, H, acc);
#pragma unroll
        for (int i = 0; i < 15; ++i) demb[i] += acc[12 + i];
    }
    __syncthreads();
}

__global__ void __launch_bounds__(kBwdThreads)
field_mlp_bwd_kernel(umhs_points P, int64_t n_total, const float *__restrict__ enc, umhs_mlp_params mp,
                     umhs_field_out_grads g, float *__restrict__ d_enc, umhs_mlp_grads dp) {
    extern __shared__ __align__(16) float smem[];
    const int K = mp.num_classes, spec = mp.specular;
    const int tid = threadIdx.x;
    const BwdSmem L = bwd_layout(K, spec);
    float *W = smem + L.w;
    float *Xs = smem + L.X, *H1s = smem + L.H1, *H2s = smem + L.H2, *D1s = smem + L.D1, *D2s = smem + L.D2;
    float *rowX = Xs + tid * LDX, *rowH1 = H1s + tid * LDH, *rowH2 = H2s + tid * LDH, *rowD1 = D1s + tid * LDH,
          *rowD2 = D2s + tid * LDH;
    const float invT = 1.0f / mp.temperature;
    const WOff oh = mlp3_offsets(K), of = mlp3_offsets(K + (spec ? 1 : 0));
    n_total = effective_n(P, n_total);
    const int64_t n_tiles = (n_total + kBwdThreads - 1) / kBwdThreads;

    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t n0 = tile * kBwdThreads;
        const int nsamp = (int)((n_total - n0) < kBwdThreads ? (n_total - n0) : kBwdThreads);
        const bool valid = tid < nsamp;
        const int64_t n = valid ? n0 + tid : n0;  // inactive threads shadow sample n0 with zero gradients
        const Point pt = eval_point(P, n);

        // ================= base MLP forward (needed for emb / relu masks)
        stage_w(W + kBaseW0, mp.base_w0, H, ENC, H, ENC, tid, kBwdThreads);
        stage_b(W + kBaseB0, mp.base_b0, H, H, tid, kBwdThreads);
        stage_w(W + kBaseW1, mp.base_w1, 16, H, 16, H, tid, kBwdThreads);
        stage_b(W + kBaseB1, mp.base_b1, 16, 16, tid, kBwdThreads);
        __syncthreads();
        float o16[16];
        {
            float in[ENC];
            const float4 *e4 = reinterpret_cast<const float4 *>(enc + enc_row(P, n) * ENC);
#pragma unroll
            for (int i = 0; i < ENC / 4; ++i) {
                const float4 t = __ldg(e4 + i);
                in[4 * i] = t.x; in[4 * i + 1] = t.y; in[4 * i + 2] = t.z; in[4 * i + 3] = t.w;
            }
            float h[H];
            dense_row<ENC, true>(W + kBaseW0, W + kBaseB0, in, rowH1, H);
            load_row<H>(rowH1, h);
            dense_row<H, false>(W + kBaseW1, W + kBaseB1, h, rowD1, 16);
            load_row<16>(rowD1, o16);
        }
        __syncthreads();

        float x[XW], pe[12], demb[15];
        pos_encoding(pt, pe);
#pragma unroll
        for (int i = 0; i < 12; ++i) x[i] = pe[i];
#pragma unroll
        for (int i = 0; i < 15; ++i) {
            x[12 + i] = o16[1 + i];
            demb[i] = (valid && g.d_embedding) ? __ldg(g.d_embedding + n * 15 + i) : 0.f;
        }
        x[27] = 0.f;

        // ================= mlp_head
        stage_mlp3(W, oh, mp.head_w0, mp.head_b0, mp.head_w1, mp.head_b1, mp.head_w2, mp.head_b2, K, tid, kBwdThreads);
        __syncthreads();
        mlp3_backward(0, K, spec, invT, valid, nsamp, n, x, W, oh, rowX, rowH1, rowH2, rowD1, rowD2, Xs, H1s, H2s, D1s,
                      D2s, g.d_scalar, nullptr, dp.head_w0, dp.head_b0, dp.head_w1, dp.head_b1, dp.head_w2, dp.head_b2,
                      demb);
        // ================= feature_mlp
        stage_mlp3(W, of, mp.feat_w0, mp.feat_b0, mp.feat_w1, mp.feat_b1, mp.feat_w2, mp.feat_b2, K + (spec ? 1 : 0),
                   tid, kBwdThreads);
        __syncthreads();
        mlp3_backward(1, K, spec, invT, valid, nsamp, n, x, W, of, rowX, rowH1, rowH2, rowD1, rowD2, Xs, H1s, H2s, D1s,
                      D2s, g.d_abund, g.d_s1, dp.feat_w0, dp.feat_b0, dp.feat_w1, dp.feat_b1, dp.feat_w2, dp.feat_b2,
                      demb);
        // ================= mlp_directional, first layer (no dgrad: its inputs carry no parameters)
        if (spec && g.d_spec_hidden) {
            stage_w(W + kDirW0, mp.dir_w0, 16, 28, 16, XW, tid, kBwdThreads);
            stage_b(W + kDirB0, mp.dir_b0, 16, 16, tid, kBwdThreads);
            __syncthreads();
            float xs[XW], sh[16];
            sh_deg3((pt.dx + 1.0f) / 2.0f, (pt.dy + 1.0f) / 2.0f, (pt.dz + 1.0f) / 2.0f, sh);
#pragma unroll
            for (int i = 0; i < 16; ++i) xs[i] = sh[i];
#pragma unroll
            for (int i = 0; i < 12; ++i) xs[16 + i] = pe[i];
#pragma unroll
            for (int i = 0; i < XW; i += 4)
                *reinterpret_cast<float4 *>(rowX + i) = make_float4(xs[i], xs[i + 1], xs[i + 2], xs[i + 3]);
            dense_row<XW, true>(W + kDirW0, W + kDirB0, xs, rowH2, 16);
#pragma unroll
            for (int i = 0; i < 16; ++i)
                rowD2[i] = (valid && rowH2[i] > 0.f) ? __ldg(g.d_spec_hidden + n * 16 + i) : 0.f;
#pragma unroll
            for (int i = 16; i < OUTP; ++i) rowD2[i] = 0.f;
            __syncthreads();
            // dWd0[j][i] = sum_s D2[s][j] X[s][i]: p = i (28 of 64 threads), q = j (16)
            wgrad_tile<8>(Xs, LDX, 28, D2s, LDH, 16, nsamp, dp.dir_w0, 1, 28, nullptr);
            colsum_tile(D2s, LDH, 16, nsamp, dp.dir_b0);
            __syncthreads();
        }
        // ================= base MLP backward
        stage_w(W + kBaseW0, mp.base_w0, H, ENC, H, ENC, tid, kBwdThreads);
        stage_b(W + kBaseB0, mp.base_b0, H, H, tid, kBwdThreads);
        stage_w(W + kBaseW1, mp.base_w1, 16, H, 16, H, tid, kBwdThreads);
        stage_b(W + kBaseB1, mp.base_b1, 16, 16, tid, kBwdThreads);
        __syncthreads();
        {
            float in[ENC];
            const float4 *e4 = reinterpret_cast<const float4 *>(enc + enc_row(P, n) * ENC);
#pragma unroll
            for (int i = 0; i < ENC / 4; ++i) {
                const float4 t = __ldg(e4 + i);
                in[4 * i] = t.x; in[4 * i + 1] = t.y; in[4 * i + 2] = t.z; in[4 * i + 3] = t.w;
                *reinterpret_cast<float4 *>(rowX + 4 * i) = t;
            }
            dense_row<ENC, true>(W + kBaseW0, W + kBaseB0, in, rowH1, H);
        }
        {
            // d out16: density = trunc_exp(o0) * sel -> d o0 = d_sigma * sel * exp(clamp(o0, -15, 15))
            const float dsg = (valid && g.d_sigma) ? __ldg(g.d_sigma + n) : 0.f;
            rowD2[0] = dsg * pt.sel * expf(fminf(fmaxf(o16[0], -15.0f), 15.0f));
#pragma unroll
            for (int i = 0; i < 15; ++i) rowD2[1 + i] = valid ? demb[i] : 0.f;
#pragma unroll
            for (int i = 16; i < OUTP; ++i) rowD2[i] = 0.f;
        }
        __syncthreads();
        // dWb1[j][i] = sum_s D2[s][j] H1[s][i]: p = i, q = j (16)
        wgrad_tile<8>(H1s, LDH, H, D2s, LDH, 16, nsamp, dp.base_w1, 1, H, nullptr);
        colsum_tile(D2s, LDH, 16, nsamp, dp.base_b1);
        {
            float acc[H];
            dgrad_row<H>(W + kBaseW1, rowD2, 16, acc);
#pragma unroll
            for (int i = 0; i < H; i += 4) {
                const float4 h = *reinterpret_cast<const float4 *>(rowH1 + i);
                *reinterpret_cast<float4 *>(rowD1 + i) = make_float4(h.x > 0.f ? acc[i] : 0.f, h.y > 0.f ? acc[i + 1] : 0.f,
                                                                     h.z > 0.f ? acc[i + 2] : 0.f, h.w > 0.f ? acc[i + 3] : 0.f);
            }
        }
        __syncthreads();
        // dWb0[j][i] = sum_s D1[s][j] X[s][i], i < 32
        wgrad_tile<16>(D1s, LDH, H, Xs, LDX, ENC, nsamp, dp.base_w0, ENC, 1, dp.base_b0);
        if (d_enc) {
            float acc[ENC];
            dgrad_row<ENC>(W + kBaseW0, rowD1, H, acc);
            if (valid) {
                float4 *dst = reinterpret_cast<float4 *>(d_enc + n * ENC);
#pragma unroll
                for (int i = 0; i < ENC / 4; ++i) dst[i] = make_float4(acc[4 * i], acc[4 * i + 1], acc[4 * i + 2], acc[4 * i + 3]);
            }
        }
        __syncthreads();
    }
}

static int check_params(const umhs_mlp_params *p, bool density_only) {
    if (!p || !p->base_w0 || !p->base_b0 || !p->base_w1 || !p->base_b1) return -1;
    if (density_only) return 0;
    if (p->num_classes < 1 || p->num_classes > UMHS_MAX_CLASSES) return -1;
    if (!(p->temperature > 0.f)) return -1;
    if (!p->head_w0 || !p->head_b0 || !p->head_w1 || !p->head_b1 || !p->head_w2 || !p->head_b2) return -1;
    if (!p->feat_w0 || !p->feat_b0 || !p->feat_w1 || !p->feat_b1 || !p->feat_w2 || !p->feat_b2) return -1;
    if (p->specular && (!p->dir_w0 || !p->dir_b0)) return -1;
    return 0;
}

static int grid_for(int64_t n, int threads, int per_sm) {
    const int64_t want = (n + threads - 1) / threads;
    const int64_t cap = (int64_t)kNumSMs * per_sm;
    return (int)(want < cap ? (want < 1 ? 1 : want) : cap);
}

}  // namespace umhs

using namespace umhs;

extern "C" int umhs_density_fwd(const umhs_points *pts, int64_t n, const float *enc, const umhs_mlp_params *p,
                                float *sigma, void *stream) {
    UMHS_CHECK_ARG(pts && enc && sigma, "null argument");
    UMHS_CHECK_ARG(check_params(p, true) == 0, "bad MLP parameters");
    if (n <= 0) return UMHS_OK;
    umhs_field_out out = {};
    out.sigma = sigma;
    const size_t bytes = (size_t)fwd_layout(1, 0, true, kFwdThreads).total * sizeof(float);
    UMHS_CHECK_CUDA(cudaFuncSetAttribute(field_mlp_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)bytes));
    field_mlp_fwd_kernel<true><<<grid_for(n, kFwdThreads, 2), kFwdThreads, bytes, as_stream(stream)>>>(*pts, n, enc, *p,
                                                                                                       out);
    UMHS_CHECK_LAUNCH();
    return UMHS_OK;
}

extern "C" int umhs_field_mlp_fwd(const umhs_points *pts, int64_t n, const float *enc, const umhs_mlp_params *p,
                                  const umhs_field_out *out, void *stream) {
    UMHS_CHECK_ARG(pts && enc && out && out->sigma && out->scalar && out->abund, "null argument");
    UMHS_CHECK_ARG(check_params(p, false) == 0, "bad MLP parameters");
    if (n <= 0) return UMHS_OK;
    const size_t bytes = (size_t)fwd_layout(p->num_classes, p->specular, false, kFwdThreads).total * sizeof(float);
    UMHS_CHECK_CUDA(cudaFuncSetAttribute(field_mlp_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)bytes));
    field_mlp_fwd_kernel<false><<<grid_for(n, kFwdThreads, 1), kFwdThreads, bytes, as_stream(stream)>>>(*pts, n, enc,
                                                                                                        *p, *out);
    UMHS_CHECK_LAUNCH();
    return UMHS_OK;
}

extern "C" int umhs_field_mlp_bwd(const umhs_points *pts, int64_t n, const float *enc, const umhs_mlp_params *p,
                                  const umhs_field_out_grads *g, float *d_enc, const umhs_mlp_grads *dp, void *stream) {
    UMHS_CHECK_ARG(pts && enc && g && dp, "null argument");
    UMHS_CHECK_ARG(check_params(p, false) == 0, "bad MLP parameters");
    UMHS_CHECK_ARG(g->d_scalar && g->d_abund, "d_scalar/d_abund are required");
    UMHS_CHECK_ARG(dp->base_w0 && dp->base_b0 && dp->base_w1 && dp->base_b1 && dp->head_w0 && dp->head_b0 &&
                       dp->head_w1 && dp->head_b1 && dp->head_w2 && dp->head_b2 && dp->feat_w0 && dp->feat_b0 &&
                       dp->feat_w1 && dp->feat_b1 && dp->feat_w2 && dp->feat_b2,
                   "null parameter-gradient pointer");
    UMHS_CHECK_ARG(!p->specular || (dp->dir_w0 && dp->dir_b0), "specular needs dir_w0/dir_b0 gradients");
    if (n <= 0) return UMHS_OK;
    const size_t bytes = (size_t)bwd_layout(p->num_classes, p->specular).total * sizeof(float);
    UMHS_CHECK_CUDA(cudaFuncSetAttribute(field_mlp_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    field_mlp_bwd_kernel<<<grid_for(n, kBwdThreads, 1), kBwdThreads, bytes, as_stream(stream)>>>(*pts, n, enc, *p, *g,
                                                                                                 d_enc, *dp);
    UMHS_CHECK_LAUNCH();
    return UMHS_OK;
}
