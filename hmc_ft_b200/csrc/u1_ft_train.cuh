// Training gradient of the field transformation: d loss / d (CNN weights) of
// field_trans_opt.py:443-480, hand-derived (no autograd).  Included at the end of u1_ft.cu (same translation unit:
// it reuses ConvP, Lat, link_terms, the handle and the first-order kernels).
//
//   loss(W) = sum_{p=2,4,6,8} || F_FT(theta_n; beta, W) - F_plain(theta_n; 1) ||_p / vol^(1/p),  theta_n = Inv_W(theta_o)
//
// With v = d loss / d (F_FT - F_plain):
//   d loss / d W = d/dW [ v . F_FT(theta_n; W) ]                      (1)  u1ft_train_force_grad
//                + (d theta_n / d W)^T [ H_FT v - H_plain v ]          (2)  u1ft_train_inverse_stage_bwd
// (1) and H_FT v are the gradient of phi(theta, W) = v . grad_theta S_FT = D_v S_FT, the directional derivative of
// S_FT along v: a forward-mode (tangent) sweep through the 8 subsets carries (theta^(k), thetadot^(k)) with
// thetadot^(0) = v, and ONE reverse sweep through that doubled computation gives d phi / d theta = H_FT v and
// d phi / d W.  Convolutions are linear, so the tangent of a layer is the same convolution applied to the tangent
// activations times GELU'; the reverse sweep carries two adjoint streams (of the pre-activations z and of their
// tangents zdot) coupled through GELU'' :  zbar = abar G'(z) + adotbar G''(z) zdot,  zdotbar = adotbar G'(z).
// (2): the reference differentiates through the UNROLLED fixed-point iterations of the inverse
// (field_trans_opt.py:245-282).  Per active link the iteration is the scalar recursion t_{i+1} = c - Delta(t_i)
// with every angle linear in t, so one thread recomputes its link's iterates and runs the adjoint recursion
// backwards; the CNN coefficients do not depend on the active links, so their adjoints of all iterations are
// summed and go through ONE first-order reverse pass of the CNN.
// Training is not the sampling hot path: the kernels here favour clarity (one thread per site, libm erfc/exp
// for the exact GELU) over speed; the batch is one chunk.
namespace u1 {

constexpr double kInvSqrt2 = 0.70710678118654752440;
constexpr double kInvSqrt2Pi = 0.39894228040143267794;

// GELU(z) = z Phi(z) (exact erf form, nn.GELU default), G' = Phi + z phi, G'' = phi (2 - z^2)
__device__ __forceinline__ void gelu3(double z, double& g, double& d1, double& d2) {
    const double Phi = 0.5 * erfc(-z * kInvSqrt2);
    const double phi = kInvSqrt2Pi * exp(-0.5 * z * z);
    g = z * Phi;
    d1 = fma(z, phi, Phi);
    d2 = phi * (2.0 - z * z);
}

enum { TR_FWD = 0, TR_FWD_FINAL = 1, TR_LIN = 2 };

// 3x3 circular cross-correlation, one thread per site, all COUT outputs in registers.
//   TR_FWD      : z = acc + b (+ add);  out = G(z), out2 = G'(z), out3 = G''(z)
//   TR_FWD_FINAL: z = acc + b;  y = G(z);  out = K = atan(y)/2pi, out2 = dK/dz, out3 = d2K/dz2
//   TR_LIN      : r = acc (+ add);  out = r * mul (+ add2)  (mul may be null);  out2 = r * mul2 (if out2)
template <int CIN, int COUT, int MODE>
__global__ void __launch_bounds__(128)
k_conv_tr(const __grid_constant__ ConvP<CIN, COUT> p, const double* __restrict__ in, const double* __restrict__ add,
          const double* __restrict__ mul, const double* __restrict__ add2, const double* __restrict__ mul2,
          double* __restrict__ out, double* __restrict__ out2, double* __restrict__ out3, int nsites, int L) {
    const int s = blockIdx.x * 128 + threadIdx.x;
    if (s >= nsites) return;
    const int V = L * L;
    const int b = s / V, r = s - b * V, x = r / L, y = r - x * L;
    const int xr[3] = {(x == 0 ? L - 1 : x - 1) * L, x * L, (x == L - 1 ? 0 : x + 1) * L};
    const int yc[3] = {(y == 0 ? L - 1 : y - 1), y, (y == L - 1 ? 0 : y + 1)};
    const double* ib = in + (size_t)b * CIN * V;
    double acc[COUT];
#pragma unroll
    for (int o = 0; o < COUT; ++o) acc[o] = 0.0;
#pragma unroll 1
    for (int c = 0; c < CIN; ++c) {
        const double* pc = ib + (size_t)c * V;
        const double* w = p.w + c * 9 * COUT;
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                const double v = pc[xr[i] + yc[j]];
#pragma unroll
                for (int o = 0; o < COUT; ++o) acc[o] = fma(w[(i * 3 + j) * COUT + o], v, acc[o]);
            }
    }
    const size_t ob = (size_t)b * COUT * V + r;
#pragma unroll
    for (int o = 0; o < COUT; ++o) {
        const size_t idx = ob + (size_t)o * V;
        if (MODE == TR_LIN) {
            double rv = acc[o];
            if (add) rv += add[idx];
            double ov = mul ? rv * mul[idx] : rv;
            if (add2) ov += add2[idx];
            out[idx] = ov;
            if (out2) out2[idx] = rv * mul2[idx];
        } else {
            double z = acc[o] + p.b[o];
            if (MODE == TR_FWD && add) z += add[idx];
            double g, d1, d2;
            gelu3(z, g, d1, d2);
            if (MODE == TR_FWD) {
                out[idx] = g; out2[idx] = d1; out3[idx] = d2;
            } else {
                const double q = 1.0 / (1.0 + g * g);
                out[idx] = atan(g) / kPi / 2;                                   // cnn_model_opt.py:46
                out2[idx] = d1 * q / kTwoPi;
                out3[idx] = (d2 * (1.0 + g * g) - 2.0 * g * d1 * d1) * q * q / kTwoPi;
            }
        }
    }
}

// Weight gradient of one layer: partial[blk][(o*CIN+c)*9+tap] = sum over the block's sites of
// zbar[o](x,y) * ain[c](x+i-1, y+j-1), partial[blk][12*CIN*9 + o] = sum zbar[o].  One thread per (o,c).
template <int CIN>
__global__ void __launch_bounds__(12 * CIN)
k_conv_wgrad(const double* __restrict__ zbar, const double* __restrict__ ain, double* __restrict__ partial,
             int nsites, int L, int sites_per_block) {
    const int o = threadIdx.x / CIN, c = threadIdx.x - o * CIN;
    const int V = L * L;
    const int s0 = blockIdx.x * sites_per_block, s1 = min(nsites, s0 + sites_per_block);
    double acc[9], bacc = 0.0;
#pragma unroll
    for (int t = 0; t < 9; ++t) acc[t] = 0.0;
    for (int s = s0; s < s1; ++s) {
        const int b = s / V, r = s - b * V, x = r / L, y = r - x * L;
        const double zb = zbar[((size_t)b * kW + o) * V + r];
        const double* pc = ain + ((size_t)b * CIN + c) * V;
        const int xr[3] = {(x == 0 ? L - 1 : x - 1) * L, x * L, (x == L - 1 ? 0 : x + 1) * L};
        const int yc[3] = {(y == 0 ? L - 1 : y - 1), y, (y == L - 1 ? 0 : y + 1)};
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) acc[i * 3 + j] = fma(zb, pc[xr[i] + yc[j]], acc[i * 3 + j]);
        bacc += zb;
    }
    double* po = partial + (size_t)blockIdx.x * (kW * CIN * 9 + kW);
#pragma unroll
    for (int t = 0; t < 9; ++t) po[(o * CIN + c) * 9 + t] = acc[t];
    if (c == 0) po[kW * CIN * 9 + o] = bacc;
}
// grad[i] += sum over blocks (fixed order) of partial[blk][i]
__global__ void k_wgrad_finish(const double* __restrict__ partial, int nblk, int n_out, double* __restrict__ grad, int with_bias) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_out) return;
    if (!with_bias && i >= n_out - kW) return;           // tangent stream: zdot = W adot has no bias term
    double s = 0.0;
    for (int k = 0; k < nblk; ++k) s += partial[(size_t)k * n_out + i];
    grad[i] += s;
}

// Tangent of the features (field_trans_opt.py:110-131) along thetadot.
__global__ void k_ft_features_tan(const double* __restrict__ th, const double* __restrict__ thd, double* __restrict__ Xd,
                                  int k, int nsites, int L) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nsites) return;
    const int V = L * L;
    const int b = s / V, r = s - b * V, x = r / L, y = r - x * L;
    const Subset ss = subset_of(k);
    Lat g{th + (size_t)b * 2 * V, th + (size_t)b * 2 * V + V, L};
    Lat gd{thd + (size_t)b * 2 * V, thd + (size_t)b * 2 * V + V, L};
    const bool active_site = ((x & 1) == ss.a) && ((y & 1) == ss.b);
    const bool mP = ss.mu == 0 ? ((x & 1) != ss.a) : ((y & 1) != ss.b);
    const bool mR = !active_site;
    double dsP = 0.0, dcP = 0.0, dsR = 0.0, dcR = 0.0;
    if (mP) { double sn, cs; sincos(g.P(x, y), &sn, &cs); const double pd = gd.P(x, y); dsP = cs * pd; dcP = -sn * pd; }
    if (mR) {
        double sn, cs;
        sincos(ss.mu == 0 ? g.R1(x, y) : g.R0(x, y), &sn, &cs);
        const double rd = ss.mu == 0 ? gd.R1(x, y) : gd.R0(x, y);
        dsR = cs * rd; dcR = -sn * rd;
    }
    double* o = Xd + (size_t)b * kXC * V + r;
    o[0] = dsP;
    o[(size_t)V] = dcP;
    o[(size_t)2 * V] = ss.mu == 0 ? 0.0 : dsR;
    o[(size_t)3 * V] = ss.mu == 0 ? dsR : 0.0;
    o[(size_t)4 * V] = ss.mu == 0 ? 0.0 : dcR;
    o[(size_t)5 * V] = ss.mu == 0 ? dcR : 0.0;
}

// thetadot^(k+1) = thetadot^(k) + Deltadot_k ;  phi partial: -Jdot/(1+J) per block (optional)
__global__ void __launch_bounds__(kUpdThreads)
k_ft_update_tan(const double* __restrict__ th, const double* __restrict__ thd, const double* __restrict__ K,
                const double* __restrict__ Kd, double* __restrict__ outd, double* __restrict__ partial, int k, int L, int nblk) {
    __shared__ double red[32];
    const int V = L * L;
    const int b = blockIdx.y;
    const int r = blockIdx.x * kUpdThreads + threadIdx.x;
    const Subset ss = subset_of(k);
    double ps[1] = {0.0};
    if (r < V) {
        const int x = r / L, y = r - x * L;
        const double* t0 = th + (size_t)b * 2 * V;
        const double* d0 = thd + (size_t)b * 2 * V;
        double o0 = d0[r], o1 = d0[V + r];
        if (((x & 1) == ss.a) && ((y & 1) == ss.b)) {
            Lat g{t0, t0 + V, L}, gd{d0, d0 + V, L};
            LinkTerms t, td;
            link_terms(g, ss.mu, x, y, t);
            link_terms(gd, ss.mu, x, y, td);
            const double* Kb = K + (size_t)b * kW * V + r;
            const double* Kdb = Kd + (size_t)b * kW * V + r;
            double dd = 0.0, J = 0.0, Jd = 0.0;
#pragma unroll
            for (int q = 0; q < 6; ++q) {
                double sn, cs;
                sincos(t.A[q], &sn, &cs);
                const size_t ch = (size_t)term_channel(ss.mu, q) * V;
                const double kq = Kb[ch], kd = Kdb[ch], sg = term_sign(ss.mu, q), ad = td.A[q];
                dd += sg * (kd * sn + kq * cs * ad);
                J -= kq * cs;
                Jd += -kd * cs + kq * sn * ad;
            }
            if (ss.mu == 0) o0 += dd; else o1 += dd;
            ps[0] = -Jd / (1.0 + J);
        }
        double* ob = outd + (size_t)b * 2 * V;
        ob[r] = o0;
        ob[V + r] = o1;
    }
    if (partial) {
        block_sum<1>(ps, red);
        if (threadIdx.x == 0) partial[(size_t)b * nblk + blockIdx.x] = ps[0];
    }
}

// Reverse of one subset's update node in the doubled (value, tangent) computation.
//   g, gd       : adjoints of theta^(k+1), thetadot^(k+1)
//   h1, Ef      : dK/dz of the final layer and (d2K/dz2) * zdot_f
//   Zv, Zt      : adjoints of the final pre-activation z_f and of its tangent (dense [B][12][V], pre-zeroed)
//   AP/AR, APd/ARd : adjoints of the update's plaquette / rectangle angles and of their tangents
__global__ void __launch_bounds__(kUpdThreads)
k_ft_update_bwd2(const double* __restrict__ th, const double* __restrict__ thd, const double* __restrict__ K,
                 const double* __restrict__ Kd, const double* __restrict__ h1, const double* __restrict__ Ef,
                 const double* __restrict__ g, const double* __restrict__ gd,
                 double* __restrict__ Zv, double* __restrict__ Zt, double* __restrict__ AP, double* __restrict__ AR,
                 double* __restrict__ APd, double* __restrict__ ARd, int k, int nsites, int L) {
    const int s = blockIdx.x * kUpdThreads + threadIdx.x;
    if (s >= nsites) return;
    const int V = L * L;
    const int b = s / V, r = s - b * V, x = r / L, y = r - x * L;
    const Subset ss = subset_of(k);
    if (((x & 1) != ss.a) || ((y & 1) != ss.b)) return;
    const double* t0 = th + (size_t)b * 2 * V;
    const double* d0 = thd + (size_t)b * 2 * V;
    Lat gl{t0, t0 + V, L}, gdl{d0, d0 + V, L};
    LinkTerms t, td;
    link_terms(gl, ss.mu, x, y, t);
    link_terms(gdl, ss.mu, x, y, td);
    const size_t base = (size_t)b * kW * V + r;
    double sn[6], cs[6], kq[6], kd[6];
    double J = 0.0, Jd = 0.0;
#pragma unroll
    for (int q = 0; q < 6; ++q) {
        sincos(t.A[q], &sn[q], &cs[q]);
        const size_t ch = (size_t)term_channel(ss.mu, q) * V;
        kq[q] = K[base + ch]; kd[q] = Kd[base + ch];
        J -= kq[q] * cs[q];
        Jd += -kd[q] * cs[q] + kq[q] * sn[q] * td.A[q];
    }
    const double a = g[(size_t)b * 2 * V + (size_t)ss.mu * V + r];      // adjoint of Delta
    const double ad = gd[(size_t)b * 2 * V + (size_t)ss.mu * V + r];    // adjoint of Deltadot
    const double inv = 1.0 / (1.0 + J);
    const double bb = -inv;                                             // d psi / d Jdot,  psi = -Jdot/(1+J)
    const double cc = Jd * inv * inv;                                   // d psi / d J
    double* APb = AP + (size_t)b * V;  double* ARb = AR + (size_t)b * V;
    double* APdb = APd + (size_t)b * V;  double* ARdb = ARd + (size_t)b * V;
#pragma unroll
    for (int q = 0; q < 6; ++q) {
        const double sg = term_sign(ss.mu, q), A_d = td.A[q];
        const size_t ch = (size_t)term_channel(ss.mu, q) * V;
        const double kbar = a * sg * sn[q] + ad * sg * cs[q] * A_d - cc * cs[q] + bb * sn[q] * A_d;
        const double kdbar = ad * sg * sn[q] - bb * cs[q];
        Zv[base + ch] = kbar * h1[base + ch] + kdbar * Ef[base + ch];
        Zt[base + ch] = kdbar * h1[base + ch];
        const double abar = a * sg * kq[q] * cs[q] + ad * sg * (kd[q] * cs[q] - kq[q] * sn[q] * A_d) + cc * kq[q] * sn[q] +
                            bb * (kd[q] * sn[q] + kq[q] * cs[q] * A_d);
        const double adbar = ad * sg * kq[q] * cs[q] + bb * kq[q] * sn[q];
        if (q < 2) { APb[t.loc[q]] = abar; APdb[t.loc[q]] = adbar; }
        else { ARb[t.loc[q]] = abar; ARdb[t.loc[q]] = adbar; }
    }
}

// Total plaquette adjoints of subset k in the doubled computation: PT (value angles) and PTd (tangent angles).
__global__ void k_ft_feat_bwd2(const double* __restrict__ th, const double* __restrict__ thd,
                               const double* __restrict__ Xb, const double* __restrict__ Xdb,
                               const double* __restrict__ AP, const double* __restrict__ AR,
                               const double* __restrict__ APd, const double* __restrict__ ARd,
                               double* __restrict__ PT, double* __restrict__ PTd, int k, int nsites, int L) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nsites) return;
    const int V = L * L;
    const int b = s / V, r = s - b * V, x = r / L, y = r - x * L;
    const Subset ss = subset_of(k);
    const double* t0 = th + (size_t)b * 2 * V;
    const double* d0 = thd + (size_t)b * 2 * V;
    Lat g{t0, t0 + V, L}, gd{d0, d0 + V, L};
    const double* xb = Xb + (size_t)b * kXC * V;
    const double* xdb = Xdb + (size_t)b * kXC * V;
    const double* ARb = AR + (size_t)b * V;
    const double* ARdb = ARd + (size_t)b * V;
    const bool mP = ss.mu == 0 ? ((x & 1) != ss.a) : ((y & 1) != ss.b);
    double pt = mP ? 0.0 : AP[(size_t)b * V + r];
    double ptd = mP ? 0.0 : APd[(size_t)b * V + r];
    if (mP) {
        double sn, cs;
        sincos(g.P(x, y), &sn, &cs);
        const double pd = gd.P(x, y);
        pt += xb[r] * cs - xb[(size_t)V + r] * sn - xdb[r] * sn * pd - xdb[(size_t)V + r] * cs * pd;
        ptd += xdb[r] * cs - xdb[(size_t)V + r] * sn;
    }
    auto cnn_rect = [&](int xx, int yy, double& rb, double& rdb) {
        const int xw = wrapi(xx, L), yw = wrapi(yy, L);
        if (((xw & 1) == ss.a) && ((yw & 1) == ss.b)) return;             // masked out
        const int q = xw * L + yw;
        const int chS = ss.mu == 0 ? 3 : 2, chC = ss.mu == 0 ? 5 : 4;
        double sn, cs;
        const double R = ss.mu == 0 ? g.R1(xw, yw) : g.R0(xw, yw);
        const double Rd = ss.mu == 0 ? gd.R1(xw, yw) : gd.R0(xw, yw);
        sincos(R, &sn, &cs);
        const double bs = xb[(size_t)chS * V + q], bc = xb[(size_t)chC * V + q];
        const double ds = xdb[(size_t)chS * V + q], dc = xdb[(size_t)chC * V + q];
        rb += bs * cs - bc * sn - ds * sn * Rd - dc * cs * Rd;
        rdb += ds * cs - dc * sn;
    };
    if (ss.mu == 0) {   // update angles are R0 (AR), CNN angles are R1
        const int q2 = g.loc(x - 1, y);
        pt += ARb[r] + ARb[q2];
        ptd += ARdb[r] + ARdb[q2];
        cnn_rect(x, y, pt, ptd);
        cnn_rect(x, y - 1, pt, ptd);
    } else {            // update angles are R1 (AR), CNN angles are R0
        const int q2 = g.loc(x, y - 1);
        pt += ARb[r] + ARb[q2];
        ptd += ARdb[r] + ARdb[q2];
        cnn_rect(x, y, pt, ptd);
        cnn_rect(x - 1, y, pt, ptd);
    }
    PT[(size_t)b * V + r] = pt;
    PTd[(size_t)b * V + r] = ptd;
}

// Hessian-vector product of the plain action: out = d/d eps grad S(theta + eps w), S = -beta sum cos P.
__global__ void k_plain_hvp(const double* __restrict__ th, const double* __restrict__ w, double* __restrict__ out,
                            int nsites, int L, double beta) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nsites) return;
    const int V = L * L;
    const int b = s / V, r = s - b * V, x = r / L, y = r - x * L;
    Lat g{th + (size_t)b * 2 * V, th + (size_t)b * 2 * V + V, L};
    Lat gw{w + (size_t)b * 2 * V, w + (size_t)b * 2 * V + V, L};
    const double c0 = cos(g.P(x, y)) * gw.P(x, y);
    const double cl = cos(g.P(x, y - 1)) * gw.P(x, y - 1);
    const double cu = cos(g.P(x - 1, y)) * gw.P(x - 1, y);
    out[(size_t)b * 2 * V + r] = beta * (c0 - cl);
    out[(size_t)b * 2 * V + V + r] = beta * (-c0 + cu);
}

// v = d loss / d d,  d = a - b,  loss = sum_p N_p^(1/p) / vol^(1/p),  N_p = sum |d|^p  (sums4 = N_2, N_4, N_6, N_8)
__global__ void k_loss_grad(const double* __restrict__ a, const double* __restrict__ b, const double* __restrict__ sums4,
                            double vol, size_t n, double* __restrict__ v) {
    double c[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const double p = 2.0 * (k + 1);
        c[k] = sums4[k] > 0.0 ? pow(sums4[k], 1.0 / p - 1.0) / pow(vol, 1.0 / p) : 0.0;
    }
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const double d = a[i] - b[i], d2 = d * d;
        v[i] = d * (c[0] + d2 * (c[1] + d2 * (c[2] + d2 * c[3])));
    }
}

constexpr int kMaxInvIter = 256;
// Adjoint of the unrolled fixed-point iteration of one inverse stage (field_trans_opt.py:245-282), per active link.
//   thc  : stage input theta_curr;  K, h1 : coefficients and dK/dz of the final layer at theta_curr
//   ubar : adjoint of the stage output;  Zv (pre-zeroed) : adjoint of z_f;  AP / AR : angle adjoints
__global__ void __launch_bounds__(kUpdThreads)
k_ft_inverse_bwd(const double* __restrict__ thc, const double* __restrict__ K, const double* __restrict__ h1,
                 const double* __restrict__ ubar, double* __restrict__ Zv, double* __restrict__ AP, double* __restrict__ AR,
                 int k, int n_iter, int nsites, int L) {
    const int s = blockIdx.x * kUpdThreads + threadIdx.x;
    if (s >= nsites) return;
    const int V = L * L;
    const int b = s / V, r = s - b * V, x = r / L, y = r - x * L;
    const Subset ss = subset_of(k);
    if (((x & 1) != ss.a) || ((y & 1) != ss.b)) return;
    const double* t0 = thc + (size_t)b * 2 * V;
    Lat gl{t0, t0 + V, L};
    LinkTerms t;
    link_terms(gl, ss.mu, x, y, t);
    const size_t base = (size_t)b * kW * V + r;
    double kq[6], sg[6];
#pragma unroll
    for (int q = 0; q < 6; ++q) { kq[q] = K[base + (size_t)term_channel(ss.mu, q) * V]; sg[q] = term_sign(ss.mu, q); }
    // every angle of the link is linear in the link itself with slope sigma_q = -sg_q:  A_q(t) = A_q(c) - sg_q (t - c)
    double ti[kMaxInvIter];
    double tc = 0.0;                                        // t - c
    for (int i = 0; i < n_iter; ++i) {
        ti[i] = tc;
        double delta = 0.0;
#pragma unroll
        for (int q = 0; q < 6; ++q) delta += kq[q] * (sg[q] * sin(t.A[q] - sg[q] * tc));
        tc = -delta;                                        // t_{i+1} = c - Delta(t_i)
    }
    double tbar = ubar[(size_t)b * 2 * V + (size_t)ss.mu * V + r];
    double kbar[6] = {0, 0, 0, 0, 0, 0}, asum[6] = {0, 0, 0, 0, 0, 0};
    for (int i = n_iter - 1; i >= 0; --i) {
        double Ji = 0.0;
#pragma unroll
        for (int q = 0; q < 6; ++q) {
            double sn, cs;
            sincos(t.A[q] - sg[q] * ti[i], &sn, &cs);
            kbar[q] -= tbar * sg[q] * sn;
            asum[q] -= tbar * sg[q] * kq[q] * cs;
            Ji -= kq[q] * cs;
        }
        tbar = -tbar * Ji;                                  // d t_{i+1} / d t_i = -J(t_i)
    }
    double* APb = AP + (size_t)b * V;
    double* ARb = AR + (size_t)b * V;
#pragma unroll
    for (int q = 0; q < 6; ++q) {
        const size_t ch = (size_t)term_channel(ss.mu, q) * V;
        Zv[base + ch] = kbar[q] * h1[base + ch];
        if (q < 2) APb[t.loc[q]] = asum[q]; else ARb[t.loc[q]] = asum[q];
    }
}

}  // namespace u1

using namespace u1;

// ---------------------------------------------------------------------------------------
// Host side of the training path
// ---------------------------------------------------------------------------------------
struct TrainWs {
    double* th[9]; double* thd[9];                 // theta^(k), thetadot^(k)         [C][2][V]
    double* X; double* Xd;                         // per subset: features / tangent  [8][C][6][V]
    double* A; double* Ad;                         // per subset, per layer output    [8][nl-1][C][12][V]
    double* D1; double* D2; double* E;             // per subset, per layer           [8][nl][C][12][V]
    double* K; double* Kd;                         // per subset                      [8][C][12][V]
    double* Zv[3]; double* Zt[3]; double* M;       // adjoint streams                 [C][12][V]
    double* Xb; double* Xdb;                       // [C][6][V]
    double* AP; double* AR; double* APd; double* ARd; double* PT; double* PTd;   // [C][V]
    double* g[2]; double* gd[2];                   // [C][2][V]
    double* partial;                               // weight-gradient partials / phi partials
    size_t f6, f12;
    int nl, nblk_w, spb;
    size_t bytes;
};

static TrainWs carve_train(void* ws, int nl, int C, int L) {
    TrainWs w{};
    const size_t V = (size_t)L * L;
    const size_t f1 = al((size_t)C * V * 8), f2 = al((size_t)C * 2 * V * 8), f6 = al((size_t)C * kXC * V * 8), f12 = al((size_t)C * kW * V * 8);
    w.f6 = f6; w.f12 = f12; w.nl = nl;
    w.spb = 256;
    w.nblk_w = (int)(((size_t)C * V + w.spb - 1) / w.spb);
    char* p = (char*)ws;
    auto take = [&](size_t n) { double* r = (double*)p; p += n; return r; };
    for (int i = 0; i < 9; ++i) { w.th[i] = take(f2); w.thd[i] = take(f2); }
    w.X = take(8 * f6); w.Xd = take(8 * f6);
    w.A = take((size_t)8 * (nl - 1) * f12); w.Ad = take((size_t)8 * (nl - 1) * f12);
    w.D1 = take((size_t)8 * nl * f12); w.D2 = take((size_t)8 * nl * f12); w.E = take((size_t)8 * nl * f12);
    w.K = take(8 * f12); w.Kd = take(8 * f12);
    for (int i = 0; i < 3; ++i) { w.Zv[i] = take(f12); w.Zt[i] = take(f12); }
    w.M = take(f12);
    w.Xb = take(f6); w.Xdb = take(f6);
    w.AP = take(f1); w.AR = take(f1); w.APd = take(f1); w.ARd = take(f1); w.PT = take(f1); w.PTd = take(f1);
    for (int i = 0; i < 2; ++i) { w.g[i] = take(f2); w.gd[i] = take(f2); }
    const size_t np = (size_t)w.nblk_w * (kW * kW * 9 + kW);
    const size_t nphi = (size_t)C * blocks_for(V, kUpdThreads);
    w.partial = take(al((np > nphi ? np : nphi) * 8));
    w.bytes = (size_t)(p - (char*)ws);
    return w;
}
static inline double* tw_X(const TrainWs& w, int k) { return (double*)((char*)w.X + (size_t)k * w.f6); }
static inline double* tw_Xd(const TrainWs& w, int k) { return (double*)((char*)w.Xd + (size_t)k * w.f6); }
static inline double* tw_A(const TrainWs& w, int k, int l) { return (double*)((char*)w.A + ((size_t)k * (w.nl - 1) + l) * w.f12); }
static inline double* tw_Ad(const TrainWs& w, int k, int l) { return (double*)((char*)w.Ad + ((size_t)k * (w.nl - 1) + l) * w.f12); }
static inline double* tw_D1(const TrainWs& w, int k, int l) { return (double*)((char*)w.D1 + ((size_t)k * w.nl + l) * w.f12); }
static inline double* tw_D2(const TrainWs& w, int k, int l) { return (double*)((char*)w.D2 + ((size_t)k * w.nl + l) * w.f12); }
static inline double* tw_E(const TrainWs& w, int k, int l) { return (double*)((char*)w.E + ((size_t)k * w.nl + l) * w.f12); }
static inline double* tw_K(const TrainWs& w, int k) { return (double*)((char*)w.K + (size_t)k * w.f12); }
static inline double* tw_Kd(const TrainWs& w, int k) { return (double*)((char*)w.Kd + (size_t)k * w.f12); }

template <int CIN, int COUT, int MODE>
static int launch_tr(const ConvP<CIN, COUT>& p, const double* in, const double* add, const double* mul, const double* add2,
                     const double* mul2, double* out, double* out2, double* out3, int ns, int L, cudaStream_t st) {
    k_conv_tr<CIN, COUT, MODE><<<blocks_for(ns, 128), 128, 0, st>>>(p, in, add, mul, add2, mul2, out, out2, out3, ns, L);
    count_launch();
    U1_CHECK_LAUNCH();
    return U1_OK;
}

// offsets of (subset k, layer l) inside the weight blob / gradient (u1ft_create layout: W[o][c][3][3] then b[o])
static inline size_t blob_offset(int nl, int k, int l) {
    const size_t first = (size_t)kW * kXC * 9 + kW, rest = (size_t)kW * kW * 9 + kW;
    return (size_t)k * (first + (size_t)(nl - 1) * rest) + (l == 0 ? 0 : first + (size_t)(l - 1) * rest);
}

// gradW(layer) += zbar (x) ain  (+ bias gradient when with_bias)
static int wgrad_layer(const TrainWs& w, const double* zbar, const double* ain, int cin, double* grad, bool with_bias,
                       int ns, int L, cudaStream_t st) {
    const int n_out = kW * cin * 9 + kW;
    if (cin == kXC) k_conv_wgrad<kXC><<<w.nblk_w, 12 * kXC, 0, st>>>(zbar, ain, w.partial, ns, L, w.spb);
    else k_conv_wgrad<kW><<<w.nblk_w, 12 * kW, 0, st>>>(zbar, ain, w.partial, ns, L, w.spb);
    count_launch();
    U1_CHECK_LAUNCH();
    k_wgrad_finish<<<blocks_for(n_out, 128), 128, 0, st>>>(w.partial, w.nblk_w, n_out, grad, with_bias ? 1 : 0);
    count_launch();
    U1_CHECK_LAUNCH();
    return U1_OK;
}

// CNN of subset k with everything the reverse sweeps need stored: X -> A[k][*], D1, D2, K (slot ks of the workspace).
static int train_cnn_forward(u1ft_handle_t h, const TrainWs& w, int k, int ks, int C, int L, cudaStream_t st) {
    const int ns = C * L * L, nl = h->n_layers;
    const ConvP<kW, kW>* fk = &h->f[(size_t)k * (nl - 1)];
    U1_TRY((launch_tr<kXC, kW, TR_FWD>(h->f0[k], tw_X(w, ks), nullptr, nullptr, nullptr, nullptr, tw_A(w, ks, 0), tw_D1(w, ks, 0),
                                      tw_D2(w, ks, 0), ns, L, st)));
    for (int l = 1; l <= nl - 2; ++l) {
        const double* skip = (l % 2 == 0) ? tw_A(w, ks, l - 2) : nullptr;            // second conv of a ResBlock adds the block input
        U1_TRY((launch_tr<kW, kW, TR_FWD>(fk[l - 1], tw_A(w, ks, l - 1), skip, nullptr, nullptr, nullptr, tw_A(w, ks, l),
                                         tw_D1(w, ks, l), tw_D2(w, ks, l), ns, L, st)));
    }
    U1_TRY((launch_tr<kW, kW, TR_FWD_FINAL>(fk[nl - 2], tw_A(w, ks, nl - 2), nullptr, nullptr, nullptr, nullptr, tw_K(w, ks),
                                           tw_D1(w, ks, nl - 1), tw_D2(w, ks, nl - 1), ns, L, st)));
    return U1_OK;
}

// Tangent of the CNN: Xd -> Ad[k][*], E[k][*] = G''(z) zdot, Kd.
static int train_cnn_tangent(u1ft_handle_t h, const TrainWs& w, int k, int C, int L, cudaStream_t st) {
    const int ns = C * L * L, nl = h->n_layers;
    const ConvP<kW, kW>* fk = &h->f[(size_t)k * (nl - 1)];
    U1_TRY((launch_tr<kXC, kW, TR_LIN>(h->f0[k], tw_Xd(w, k), nullptr, tw_D1(w, k, 0), nullptr, tw_D2(w, k, 0), tw_Ad(w, k, 0),
                                      tw_E(w, k, 0), nullptr, ns, L, st)));
    for (int l = 1; l <= nl - 2; ++l) {
        const double* skip = (l % 2 == 0) ? tw_Ad(w, k, l - 2) : nullptr;
        U1_TRY((launch_tr<kW, kW, TR_LIN>(fk[l - 1], tw_Ad(w, k, l - 1), skip, tw_D1(w, k, l), nullptr, tw_D2(w, k, l),
                                         tw_Ad(w, k, l), tw_E(w, k, l), nullptr, ns, L, st)));
    }
    U1_TRY((launch_tr<kW, kW, TR_LIN>(fk[nl - 2], tw_Ad(w, k, nl - 2), nullptr, tw_D1(w, k, nl - 1), nullptr, tw_D2(w, k, nl - 1),
                                     tw_Kd(w, k), tw_E(w, k, nl - 1), nullptr, ns, L, st)));
    return U1_OK;
}

// Reverse of the CNN of subset k (storage slot ks).  On entry Zv[0] (and Zt[0] when with_tan) hold the adjoints of the
// final pre-activation (and of its tangent); on exit Xb (and Xdb) hold the feature adjoints.  Accumulates the weight
// gradients of the subset into gradW.
static int train_cnn_backward(u1ft_handle_t h, const TrainWs& w, int k, int ks, bool with_tan, double* gradW, int C, int L,
                              cudaStream_t st) {
    const int ns = C * L * L, nl = h->n_layers;
    const ConvP<kW, kW>* bk = &h->bt[(size_t)k * (nl - 1)];
    // z(l) lives in buffer slot[l % 3]: the skip adjoint of layer l is needed when layer l-1 produces z(l-2)
    auto ZV = [&](int l) { return w.Zv[(nl - 1 - l) % 3]; };
    auto ZT = [&](int l) { return w.Zt[(nl - 1 - l) % 3]; };
    for (int l = nl - 1; l >= 0; --l) {
        const double* ain = (l == 0) ? tw_X(w, ks) : tw_A(w, ks, l - 1);
        const int cin = (l == 0) ? kXC : kW;
        double* gl = gradW + blob_offset(nl, k, l);
        U1_TRY(wgrad_layer(w, ZV(l), ain, cin, gl, true, ns, L, st));
        if (with_tan) {
            const double* aind = (l == 0) ? tw_Xd(w, ks) : tw_Ad(w, ks, l - 1);
            U1_TRY(wgrad_layer(w, ZT(l), aind, cin, gl, false, ns, L, st));
        }
        if (l == 0) break;
        const int m = l - 1;                          // layer below: z(m) = (convT_l z(l) + skip) * G'(z_m) [+ mixing]
        // out[m] also feeds the skip of layer m+2 when m is even and m+2 is a ResBlock's second conv (<= nl-2)
        const bool skip = (m % 2 == 0) && (m + 2 <= nl - 2);
        if (with_tan) {
            U1_TRY((launch_tr<kW, kW, TR_LIN>(bk[l - 1], ZT(l), skip ? ZT(m + 2) : nullptr, tw_D1(w, ks, m), nullptr, tw_E(w, ks, m),
                                             ZT(m), w.M, nullptr, ns, L, st)));
            U1_TRY((launch_tr<kW, kW, TR_LIN>(bk[l - 1], ZV(l), skip ? ZV(m + 2) : nullptr, tw_D1(w, ks, m), w.M, nullptr,
                                             ZV(m), nullptr, nullptr, ns, L, st)));
        } else {
            U1_TRY((launch_tr<kW, kW, TR_LIN>(bk[l - 1], ZV(l), skip ? ZV(m + 2) : nullptr, tw_D1(w, ks, m), nullptr, nullptr,
                                             ZV(m), nullptr, nullptr, ns, L, st)));
        }
    }
    U1_TRY((launch_tr<kW, kXC, TR_LIN>(h->b0[k], ZV(0), nullptr, nullptr, nullptr, nullptr, w.Xb, nullptr, nullptr, ns, L, st)));
    if (with_tan)
        U1_TRY((launch_tr<kW, kXC, TR_LIN>(h->b0[k], ZT(0), nullptr, nullptr, nullptr, nullptr, w.Xdb, nullptr, nullptr, ns, L, st)));
    return U1_OK;
}

extern "C" {

size_t u1ft_train_workspace_bytes(u1ft_handle_t h, int B, int L) {
    if (!handle_ok(h) || B <= 0 || L <= 0) return 0;
    return carve_train(nullptr, h->n_layers, B, L).bytes;
}

int u1_force_hvp(const double* theta, const double* w, double* out, int B, int L, double beta, u1_stream_t stream) {
    g_launches = 0;
    if (!theta || !w || !out || B <= 0 || L <= 0) return U1_ERR_ARG;
    const int ns = B * L * L;
    k_plain_hvp<<<blocks_for(ns, 256), 256, 0, (cudaStream_t)stream>>>(theta, w, out, ns, L, beta);
    count_launch();
    U1_CHECK_LAUNCH();
    return U1_OK;
}

int u1_loss_grad(const double* a, const double* b, const double* sums4, double vol, size_t n, double* v, u1_stream_t stream) {
    g_launches = 0;
    if (!a || !b || !sums4 || !v || n == 0 || !(vol > 0)) return U1_ERR_ARG;
    k_loss_grad<<<ew_grid_ft(n), 256, 0, (cudaStream_t)stream>>>(a, b, sums4, vol, n, v);
    count_launch();
    U1_CHECK_LAUNCH();
    return U1_OK;
}

int u1ft_train_force_grad(u1ft_handle_t h, const double* theta, const double* v, int B, int L, double beta,
                          double* phi, double* Hv, double* force_check, double* gradW, size_t n_weights,
                          void* ws, size_t ws_bytes, u1_stream_t stream) {
    g_launches = 0;
    if (!handle_ok(h) || !handle_on_current_device(h)) return U1_ERR_HANDLE;
    if (!theta || !v || !Hv || !gradW || B <= 0 || L <= 0) return U1_ERR_ARG;
    if (L & 1) return U1_ERR_SHAPE;
    if (n_weights != u1ft_weight_count(h->n_res) || B > kMaxGridY) return U1_ERR_ARG;
    const int nl = h->n_layers;
    U1_TRY(check_ws_ft(ws, ws_bytes, carve_train(nullptr, nl, B, L).bytes));
    TrainWs w = carve_train(ws, nl, B, L);
    cudaStream_t st = (cudaStream_t)stream;
    const int C = B, V = L * L, ns = C * V, nblk = blocks_for(V, kUpdThreads);
    const size_t per = (size_t)2 * V;
    U1_TRY(copy_async(w.th[0], theta, (size_t)C * per * 8, st));
    U1_TRY(copy_async(w.thd[0], v, (size_t)C * per * 8, st));
    // ---- forward sweep with tangents ----
    for (int k = 0; k < 8; ++k) {
        k_ft_features<<<blocks_for(ns, 256), 256, 0, st>>>(w.th[k], tw_X(w, k), k, ns, L);
        count_launch();
        U1_CHECK_LAUNCH();
        U1_TRY(train_cnn_forward(h, w, k, k, C, L, st));
        k_ft_update<<<dim3(nblk, C), kUpdThreads, 0, st>>>(w.th[k], tw_K(w, k), w.th[k + 1], nullptr, k, L, nblk, 0);
        count_launch();
        U1_CHECK_LAUNCH();
        k_ft_features_tan<<<blocks_for(ns, 256), 256, 0, st>>>(w.th[k], w.thd[k], tw_Xd(w, k), k, ns, L);
        count_launch();
        U1_CHECK_LAUNCH();
        U1_TRY(train_cnn_tangent(h, w, k, C, L, st));
        k_ft_update_tan<<<dim3(nblk, C), kUpdThreads, 0, st>>>(w.th[k], w.thd[k], tw_K(w, k), tw_Kd(w, k), w.thd[k + 1], nullptr,
                                                               k, L, nblk);
        count_launch();
        U1_CHECK_LAUNCH();
    }
    (void)phi;
    // ---- reverse sweep: adjoints of phi = grad S(theta8) . thetadot8 - sum Jdot/(1+J) ----
    double* g = w.g[0];  double* gn = w.g[1];
    double* gd = w.gd[0];  double* gdn = w.gd[1];
    k_plain_hvp<<<blocks_for(ns, 256), 256, 0, st>>>(w.th[8], w.thd[8], g, ns, L, beta);      // d phi / d theta8
    count_launch();
    U1_CHECK_LAUNCH();
    k_plain_force<<<blocks_for(ns, 256), 256, 0, st>>>(w.th[8], gd, ns, L, beta);              // d phi / d thetadot8
    count_launch();
    U1_CHECK_LAUNCH();
    for (int k = 7; k >= 0; --k) {
        cudaError_t e = cudaMemsetAsync(w.Zv[0], 0, (size_t)C * kW * V * 8, st);
        if (e == cudaSuccess) e = cudaMemsetAsync(w.Zt[0], 0, (size_t)C * kW * V * 8, st);
        if (e != cudaSuccess) return (int)e;
        k_ft_update_bwd2<<<blocks_for(ns, kUpdThreads), kUpdThreads, 0, st>>>(
            w.th[k], w.thd[k], tw_K(w, k), tw_Kd(w, k), tw_D1(w, k, nl - 1), tw_E(w, k, nl - 1), g, gd, w.Zv[0], w.Zt[0],
            w.AP, w.AR, w.APd, w.ARd, k, ns, L);
        count_launch();
        U1_CHECK_LAUNCH();
        U1_TRY(train_cnn_backward(h, w, k, k, true, gradW, C, L, st));
        k_ft_feat_bwd2<<<blocks_for(ns, 256), 256, 0, st>>>(w.th[k], w.thd[k], w.Xb, w.Xdb, w.AP, w.AR, w.APd, w.ARd, w.PT, w.PTd,
                                                            k, ns, L);
        count_launch();
        U1_CHECK_LAUNCH();
        k_ft_scatter<<<blocks_for(ns, 256), 256, 0, st>>>(g, w.PT, gn, ns, L);
        count_launch();
        U1_CHECK_LAUNCH();
        k_ft_scatter<<<blocks_for(ns, 256), 256, 0, st>>>(gd, w.PTd, gdn, ns, L);
        count_launch();
        U1_CHECK_LAUNCH();
        double* t = g; g = gn; gn = t;
        t = gd; gd = gdn; gdn = t;
    }
    U1_TRY(copy_async(Hv, g, (size_t)C * per * 8, st));
    if (force_check) U1_TRY(copy_async(force_check, gd, (size_t)C * per * 8, st));   // d phi / d v = F_FT(theta): self check
    return U1_OK;
}

int u1ft_train_inverse_stage_bwd(u1ft_handle_t h, int index, const double* theta_curr, int n_iter, const double* ubar_in,
                                 double* ubar_out, double* gradW, size_t n_weights, int B, int L,
                                 void* ws, size_t ws_bytes, u1_stream_t stream) {
    g_launches = 0;
    if (!handle_ok(h) || !handle_on_current_device(h)) return U1_ERR_HANDLE;
    if (!theta_curr || !ubar_in || !ubar_out || !gradW || B <= 0 || L <= 0 || index < 0 || index > 7) return U1_ERR_ARG;
    if (L & 1) return U1_ERR_SHAPE;
    if (n_iter < 0 || n_iter > kMaxInvIter || n_weights != u1ft_weight_count(h->n_res) || ubar_in == ubar_out) return U1_ERR_ARG;
    const int nl = h->n_layers;
    U1_TRY(check_ws_ft(ws, ws_bytes, carve_train(nullptr, nl, B, L).bytes));
    TrainWs w = carve_train(ws, nl, B, L);
    cudaStream_t st = (cudaStream_t)stream;
    const int C = B, V = L * L, ns = C * V;
    if (n_iter == 0)          // the stage did not converge and was skipped (field_trans_opt.py:272-276): identity
        return copy_async(ubar_out, ubar_in, (size_t)C * 2 * V * 8, st);
    const int ks = 0;                                            // storage slot
    k_ft_features<<<blocks_for(ns, 256), 256, 0, st>>>(theta_curr, tw_X(w, ks), index, ns, L);
    count_launch();
    U1_CHECK_LAUNCH();
    U1_TRY(train_cnn_forward(h, w, index, ks, C, L, st));
    cudaError_t e = cudaMemsetAsync(w.Zv[0], 0, (size_t)C * kW * V * 8, st);
    if (e != cudaSuccess) return (int)e;
    k_ft_inverse_bwd<<<blocks_for(ns, kUpdThreads), kUpdThreads, 0, st>>>(theta_curr, tw_K(w, ks), tw_D1(w, ks, nl - 1), ubar_in,
                                                                         w.Zv[0], w.AP, w.AR, index, n_iter, ns, L);
    count_launch();
    U1_CHECK_LAUNCH();
    U1_TRY(train_cnn_backward(h, w, index, ks, false, gradW, C, L, st));
    k_ft_feat_bwd<<<blocks_for(ns, 256), 256, 0, st>>>(theta_curr, w.Xb, w.AP, w.AR, w.PT, index, ns, L);
    count_launch();
    U1_CHECK_LAUNCH();
    k_ft_scatter<<<blocks_for(ns, 256), 256, 0, st>>>(ubar_in, w.PT, ubar_out, ns, L);
    count_launch();
    U1_CHECK_LAUNCH();
    return U1_OK;
}

}  // extern "C"
