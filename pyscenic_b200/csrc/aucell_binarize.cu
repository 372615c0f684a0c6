// aucell_binarize.cu -- binarization thresholds of an AUC matrix, regulon-parallel on the GPU.
//
// Replaces pyscenic.binarization.derive_threshold / binarize (reference src/pyscenic/binarization.py:12-98), which per
// regulon (a) decides whether the AUC distribution is bimodal -- Hartigan's dip test (the `diptest` package) or a BIC
// comparison of one- and two-component Gaussian mixtures --, (b) if not, takes mean + 2 std, (c) if so, fits a
// two-component mixture (sklearn EM) and minimises the Gaussian kernel density (scipy gaussian_kde + bounded Brent)
// between the two component means.  The reference does this one regulon at a time in a process pool; here every
// regulon is a row of a [regulons x cells] matrix of SORTED values and the steps are kernels over all rows:
//
//   binarize_stats_kernel   mean, population / sample variance                       (one CTA per regulon)
//   binarize_dip_kernel     Hartigan & Hartigan's dip statistic (AS 217, the greatest-convex-minorant / least-concave-
//                           majorant iteration as in Maechler's C translation); inherently sequential: one thread per
//                           regulon, the regulons run side by side
//   binarize_gmm2_kernel    EM of a two-component 1-D mixture in fp64 with sklearn's update formulas (reg_covar, the
//                           10 * eps in nk, convergence on the change of the mean log-likelihood)  (one CTA per regulon)
//   binarize_trough_kernel  scipy's bounded Brent minimiser run by one CTA per regulon; every function evaluation is
//                           a block-wide fp64 reduction of the kernel density over the regulon's cells
//
// All arithmetic is fp64.  sm_100a only; no CPU fallback.
#include "launch.cuh"

using namespace aucell_host;

namespace {

constexpr int kBT = 512;

__device__ __forceinline__ double block_sum(double v, double* s_red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();                       // s_red may still be read from the previous call
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
    for (int w = 0; w < kBT / 32; ++w) t += s_red[w];      // every thread the same order: the same sum
    return t;
}

// out[r] = { mean, variance with ddof 0, variance with ddof 1, unused }
__global__ void __launch_bounds__(kBT) binarize_stats_kernel(const double* __restrict__ cols, int64_t n, double* __restrict__ out) {
    __shared__ double s_red[kBT / 32];
    const double* x = cols + (int64_t)blockIdx.x * n;
    double s = 0.0;
    for (int64_t i = threadIdx.x; i < n; i += kBT) s += x[i];
    const double mean = block_sum(s, s_red) / (double)n;
    double q = 0.0;
    for (int64_t i = threadIdx.x; i < n; i += kBT) { const double d = x[i] - mean; q += d * d; }
    const double ss = block_sum(q, s_red);
    if (threadIdx.x == 0) {
        double* o = out + (int64_t)blockIdx.x * 4;
        o[0] = mean;
        o[1] = ss / (double)n;
        o[2] = n > 1 ? ss / (double)(n - 1) : 0.0;
        o[3] = 0.0;
    }
}

// Dip statistic of one sorted row per thread.  scratch: 4 int32 arrays of n + 1 per regulon (mn, mj, gcm, lcm), 1-based
// like the published algorithm.
__global__ void binarize_dip_kernel(const double* __restrict__ cols, int n_reg, int n, int* __restrict__ scratch,
                                    double* __restrict__ out_dip) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_reg) return;
    const double* x = cols + (int64_t)r * n - 1;           // x[1..n]
    int* mn = scratch + (int64_t)r * 4 * (n + 1);
    int* mj = mn + (n + 1);
    int* gcm = mj + (n + 1);
    int* lcm = gcm + (n + 1);
    if (n < 2 || x[n] == x[1]) {
        out_dip[r] = n < 1 ? 0.0 : 1.0 / (2.0 * n);        // (diptest with allow_zero=False: the minimal dip)
        return;
    }
    int low = 1, high = n;
    double dip = 1.0;                                      // in units of 1 / (2 n)
    // indices over which combination is necessary for the convex minorant fit
    mn[1] = 1;
    for (int j = 2; j <= n; ++j) {
        mn[j] = j - 1;
        for (;;) {
            const int mnj = mn[j], mnmnj = mn[mnj];
            if (mnj == 1 || (x[j] - x[mnj]) * (double)(mnj - mnmnj) < (x[mnj] - x[mnmnj]) * (double)(j - mnj)) break;
            mn[j] = mnmnj;
        }
    }
    // ... and for the concave majorant
    mj[n] = n;
    for (int k = n - 1; k >= 1; --k) {
        mj[k] = k + 1;
        for (;;) {
            const int mjk = mj[k], mjmjk = mj[mjk];
            if (mjk == n || (x[k] - x[mjk]) * (double)(mjk - mjmjk) < (x[mjk] - x[mjmjk]) * (double)(k - mjk)) break;
            mj[k] = mjmjk;
        }
    }
    for (;;) {
        // change points of the GCM from high to low, of the LCM from low to high
        gcm[1] = high;
        int i = 1;
        while (gcm[i] > low) { gcm[i + 1] = mn[gcm[i]]; ++i; }
        int ig = i;
        const int l_gcm = i;
        int ix = ig - 1;
        lcm[1] = low;
        i = 1;
        while (lcm[i] < high) { lcm[i + 1] = mj[lcm[i]]; ++i; }
        int ih = i;
        const int l_lcm = i;
        int iv = 2;
        // the largest distance greater than `dip` between the GCM and the LCM from low to high
        double d = 0.0;
        if (l_gcm != 2 || l_lcm != 2) {
            do {
                const int gcmix = gcm[ix], lcmiv = lcm[iv];
                if (gcmix > lcmiv) {       // the next point of either curve is from the LCM
                    const int gcmi1 = gcm[ix + 1];
                    const double dx = (double)(lcmiv - gcmi1 + 1) -
                                      (x[lcmiv] - x[gcmi1]) * (double)(gcmix - gcmi1) / (x[gcmix] - x[gcmi1]);
                    ++iv;
                    if (dx >= d) { d = dx; ig = ix + 1; ih = iv - 1; }
                } else {                   // ... from the GCM
                    const int lcmiv1 = lcm[iv - 1];
                    const double dx = (x[gcmix] - x[lcmiv1]) * (double)(lcmiv - lcmiv1) / (x[lcmiv] - x[lcmiv1]) -
                                      (double)(gcmix - lcmiv1 - 1);
                    --ix;
                    if (dx >= d) { d = dx; ig = ix + 1; ih = iv; }
                }
                if (ix < 1) ix = 1;
                if (iv > l_lcm) iv = l_lcm;
            } while (gcm[ix] != lcm[iv]);
        } else {
            d = 1.0;
        }
        if (d < dip) break;
        // the dips for the current low and high: convex minorant ...
        double dip_l = 0.0;
        for (int j = ig; j < l_gcm; ++j) {
            double max_t = 1.0;
            const int jb = gcm[j + 1], je = gcm[j];
            if (je - jb > 1 && x[je] != x[jb]) {
                const double C = (double)(je - jb) / (x[je] - x[jb]);
                for (int jj = jb; jj <= je; ++jj) {
                    const double t = (double)(jj - jb + 1) - (x[jj] - x[jb]) * C;
                    if (max_t < t) max_t = t;
                }
            }
            if (dip_l < max_t) dip_l = max_t;
        }
        // ... and concave majorant
        double dip_u = 0.0;
        for (int j = ih; j < l_lcm; ++j) {
            double max_t = 1.0;
            const int jb = lcm[j], je = lcm[j + 1];
            if (je - jb > 1 && x[je] != x[jb]) {
                const double C = (double)(je - jb) / (x[je] - x[jb]);
                for (int jj = jb; jj <= je; ++jj) {
                    const double t = (x[jj] - x[jb]) * C - (double)(jj - jb - 1);
                    if (max_t < t) max_t = t;
                }
            }
            if (dip_u < max_t) dip_u = max_t;
        }
        const double dipnew = dip_l > dip_u ? dip_l : dip_u;
        if (dip < dipnew) dip = dipnew;
        if (low == gcm[ig] && high == lcm[ih]) break;      // no improvement in low or high
        low = gcm[ig];
        high = lcm[ih];
    }
    out_dip[r] = dip / (2.0 * n);
}

// Two-component EM.  split[r]: the first `split` (sorted) values start in component 0, the rest in component 1.
// out[r] = { w0, w1, m0, m1, v0, v1, mean log-likelihood under the final parameters, iterations, converged }
__global__ void __launch_bounds__(kBT) binarize_gmm2_kernel(const double* __restrict__ cols, int64_t n,
                                                            const int* __restrict__ split, int max_iter, double tol,
                                                            double reg_covar, double* __restrict__ out) {
    __shared__ double s_red[kBT / 32];
    const double* x = cols + (int64_t)blockIdx.x * n;
    const double eps10 = 10.0 * 2.220446049250313e-16;
    const double log2pi = 1.8378770664093453;
    double w[2], m[2], v[2];
    {   // M step from the initial hard assignment
        const int64_t sp = split[blockIdx.x];
        double c0 = 0.0, s0 = 0.0, s1 = 0.0;
        for (int64_t i = threadIdx.x; i < n; i += kBT) {
            const double xi = x[i];
            if (i < sp) { c0 += 1.0; s0 += xi; } else s1 += xi;
        }
        c0 = block_sum(c0, s_red); s0 = block_sum(s0, s_red); s1 = block_sum(s1, s_red);
        const double nk0 = c0 + eps10, nk1 = ((double)n - c0) + eps10;
        m[0] = s0 / nk0; m[1] = s1 / nk1;
        double q0 = 0.0, q1 = 0.0;
        for (int64_t i = threadIdx.x; i < n; i += kBT) {
            const double xi = x[i];
            if (i < sp) { const double d = xi - m[0]; q0 += d * d; } else { const double d = xi - m[1]; q1 += d * d; }
        }
        q0 = block_sum(q0, s_red); q1 = block_sum(q1, s_red);
        v[0] = q0 / nk0 + reg_covar; v[1] = q1 / nk1 + reg_covar;
        w[0] = nk0 / (double)n; w[1] = nk1 / (double)n;
    }
    double lower = -INFINITY;
    int iters = 0, converged = 0;
    for (int it = 1; it <= max_iter; ++it) {
        iters = it;
        const double prev = lower;
        const double lw0 = log(w[0]) - 0.5 * (log2pi + log(v[0])), lw1 = log(w[1]) - 0.5 * (log2pi + log(v[1]));
        const double iv0 = 1.0 / v[0], iv1 = 1.0 / v[1];
        // E step + the sums of the M step that do not need the new means
        double ll = 0.0, n0 = 0.0, n1 = 0.0, s0 = 0.0, s1 = 0.0;
        for (int64_t i = threadIdx.x; i < n; i += kBT) {
            const double xi = x[i];
            const double d0 = xi - m[0], d1 = xi - m[1];
            const double l0 = lw0 - 0.5 * d0 * d0 * iv0, l1 = lw1 - 0.5 * d1 * d1 * iv1;
            const double mx = l0 > l1 ? l0 : l1;
            const double lse = mx + log(exp(l0 - mx) + exp(l1 - mx));
            const double r0 = exp(l0 - lse), r1 = exp(l1 - lse);
            ll += lse; n0 += r0; n1 += r1; s0 += r0 * xi; s1 += r1 * xi;
        }
        ll = block_sum(ll, s_red); n0 = block_sum(n0, s_red); n1 = block_sum(n1, s_red);
        s0 = block_sum(s0, s_red); s1 = block_sum(s1, s_red);
        const double nk0 = n0 + eps10, nk1 = n1 + eps10;
        const double nm0 = s0 / nk0, nm1 = s1 / nk1;
        // covariances around the NEW means, responsibilities from the OLD parameters
        double q0 = 0.0, q1 = 0.0;
        for (int64_t i = threadIdx.x; i < n; i += kBT) {
            const double xi = x[i];
            const double d0 = xi - m[0], d1 = xi - m[1];
            const double l0 = lw0 - 0.5 * d0 * d0 * iv0, l1 = lw1 - 0.5 * d1 * d1 * iv1;
            const double mx = l0 > l1 ? l0 : l1;
            const double lse = mx + log(exp(l0 - mx) + exp(l1 - mx));
            const double e0 = xi - nm0, e1 = xi - nm1;
            q0 += exp(l0 - lse) * e0 * e0; q1 += exp(l1 - lse) * e1 * e1;
        }
        q0 = block_sum(q0, s_red); q1 = block_sum(q1, s_red);
        m[0] = nm0; m[1] = nm1;
        v[0] = q0 / nk0 + reg_covar; v[1] = q1 / nk1 + reg_covar;
        w[0] = nk0 / (nk0 + nk1); w[1] = nk1 / (nk0 + nk1);
        lower = ll / (double)n;
        if (fabs(lower - prev) < tol) { converged = 1; break; }
    }
    // mean log-likelihood under the final parameters (GaussianMixture.score, for the BIC)
    double ll = 0.0;
    {
        const double lw0 = log(w[0]) - 0.5 * (log2pi + log(v[0])), lw1 = log(w[1]) - 0.5 * (log2pi + log(v[1]));
        const double iv0 = 1.0 / v[0], iv1 = 1.0 / v[1];
        for (int64_t i = threadIdx.x; i < n; i += kBT) {
            const double xi = x[i];
            const double d0 = xi - m[0], d1 = xi - m[1];
            const double l0 = lw0 - 0.5 * d0 * d0 * iv0, l1 = lw1 - 0.5 * d1 * d1 * iv1;
            const double mx = l0 > l1 ? l0 : l1;
            ll += mx + log(exp(l0 - mx) + exp(l1 - mx));
        }
        ll = block_sum(ll, s_red);
    }
    if (threadIdx.x == 0) {
        double* o = out + (int64_t)blockIdx.x * 9;
        o[0] = w[0]; o[1] = w[1]; o[2] = m[0]; o[3] = m[1]; o[4] = v[0]; o[5] = v[1];
        o[6] = ll / (double)n; o[7] = (double)iters; o[8] = (double)converged;
    }
}

// gaussian_kde(data)(x) for scalar x, every thread of the block gets the value
__device__ __forceinline__ double kde_at(const double* __restrict__ x, int64_t n, double at, double inv2h2, double norm,
                                         double* s_red) {
    double s = 0.0;
    for (int64_t i = threadIdx.x; i < n; i += kBT) {
        const double d = at - x[i];
        s += exp(-d * d * inv2h2);
    }
    return block_sum(s, s_red) * norm;
}

// scipy.optimize._minimize_scalar_bounded on f = gaussian_kde(row) between lo[r] and hi[r]; kde_var[r] = the kernel's
// variance (sample variance * n^(-2/5), Scott's rule as scipy applies it)
__global__ void __launch_bounds__(kBT) binarize_trough_kernel(const double* __restrict__ cols, int64_t n,
                                                              const double* __restrict__ lo, const double* __restrict__ hi,
                                                              const double* __restrict__ kde_var, double xatol, int maxiter,
                                                              double* __restrict__ out_x) {
    __shared__ double s_red[kBT / 32];
    const double* x = cols + (int64_t)blockIdx.x * n;
    const double var = kde_var[blockIdx.x];
    const double inv2h2 = 0.5 / var, norm = 1.0 / ((double)n * sqrt(2.0 * 3.141592653589793 * var));
    auto f = [&](double at) { return kde_at(x, n, at, inv2h2, norm, s_red); };
    const double sqrt_eps = sqrt(2.2e-16), golden_mean = 0.5 * (3.0 - sqrt(5.0));
    double a = lo[blockIdx.x], b = hi[blockIdx.x];
    double fulc = a + golden_mean * (b - a), nfc = fulc, xf = fulc, rat = 0.0, e = 0.0, xx = xf;
    double fx = f(xx);
    int num = 1;
    double fu = INFINITY, ffulc = fx, fnfc = fx;
    double xm = 0.5 * (a + b), tol1 = sqrt_eps * fabs(xf) + xatol / 3.0, tol2 = 2.0 * tol1;
    while (fabs(xf - xm) > (tol2 - 0.5 * (b - a))) {
        bool golden = true;
        if (fabs(e) > tol1) {              // parabolic fit?
            golden = false;
            double r = (xf - nfc) * (fx - ffulc);
            double q = (xf - fulc) * (fx - fnfc);
            double p = (xf - fulc) * q - (xf - nfc) * r;
            q = 2.0 * (q - r);
            if (q > 0.0) p = -p;
            q = fabs(q);
            r = e;
            e = rat;
            if (fabs(p) < fabs(0.5 * q * r) && p > q * (a - xf) && p < q * (b - xf)) {
                rat = (p + 0.0) / q;
                xx = xf + rat;
                if ((xx - a) < tol2 || (b - xx) < tol2) {
                    const double dm = xm - xf;
                    const double si = (dm > 0.0 ? 1.0 : (dm < 0.0 ? -1.0 : 0.0)) + (dm == 0.0 ? 1.0 : 0.0);
                    rat = tol1 * si;
                }
            } else {
                golden = true;
            }
        }
        if (golden) {
            e = xf >= xm ? a - xf : b - xf;
            rat = golden_mean * e;
        }
        const double si = (rat > 0.0 ? 1.0 : (rat < 0.0 ? -1.0 : 0.0)) + (rat == 0.0 ? 1.0 : 0.0);
        xx = xf + si * fmax(fabs(rat), tol1);
        fu = f(xx);
        ++num;
        if (fu <= fx) {
            if (xx >= xf) a = xf; else b = xf;
            fulc = nfc; ffulc = fnfc;
            nfc = xf; fnfc = fx;
            xf = xx; fx = fu;
        } else {
            if (xx < xf) a = xx; else b = xx;
            if (fu <= fnfc || nfc == xf) {
                fulc = nfc; ffulc = fnfc;
                nfc = xx; fnfc = fu;
            } else if (fu <= ffulc || fulc == xf || fulc == nfc) {
                fulc = xx; ffulc = fu;
            }
        }
        xm = 0.5 * (a + b);
        tol1 = sqrt_eps * fabs(xf) + xatol / 3.0;
        tol2 = 2.0 * tol1;
        if (num >= maxiter) break;
    }
    if (threadIdx.x == 0) out_x[blockIdx.x] = xf;
}

int check_common(int device, const void* a, const void* b, int n_reg, int64_t n_cells) {
    if (!a || !b) return fail(AUCELL_ERR_INVALID, "NULL pointer");
    if (n_reg <= 0 || n_cells <= 0) return fail(AUCELL_ERR_INVALID, "binarize: need n_regulons, n_cells > 0");
    if (n_cells > INT32_MAX - 2) return fail(AUCELL_ERR_UNSUPPORTED, "binarize: more than 2^31 cells");
    (void)device;
    return AUCELL_OK;
}

}  // namespace

extern "C" {

int aucell_binarize_stats_f64(int device, const double* d_cols, int32_t n_regulons, int64_t n_cells, double* d_out4,
                              void* stream) {
    if (int rc = check_common(device, d_cols, d_out4, n_regulons, n_cells)) return rc;
    DeviceGuard dg(device);
    if (!dg.ok) return fail(AUCELL_ERR_CUDA, "cudaSetDevice failed");
    binarize_stats_kernel<<<n_regulons, kBT, 0, (cudaStream_t)stream>>>(d_cols, n_cells, d_out4);
    launch_counter().fetch_add(1);
    CK(cudaGetLastError());
    return AUCELL_OK;
}

int aucell_binarize_dip_f64(int device, const double* d_sorted_cols, int32_t n_regulons, int64_t n_cells,
                            int32_t* d_scratch, double* d_out_dip, void* stream) {
    if (int rc = check_common(device, d_sorted_cols, d_out_dip, n_regulons, n_cells)) return rc;
    if (!d_scratch) return fail(AUCELL_ERR_INVALID, "NULL scratch");
    DeviceGuard dg(device);
    if (!dg.ok) return fail(AUCELL_ERR_CUDA, "cudaSetDevice failed");
    binarize_dip_kernel<<<(n_regulons + 31) / 32, 32, 0, (cudaStream_t)stream>>>(d_sorted_cols, n_regulons, (int)n_cells,
                                                                                 d_scratch, d_out_dip);
    launch_counter().fetch_add(1);
    CK(cudaGetLastError());
    return AUCELL_OK;
}

int aucell_binarize_gmm2_f64(int device, const double* d_sorted_cols, int32_t n_regulons, int64_t n_cells,
                             const int32_t* d_split, int32_t max_iter, double tol, double reg_covar, double* d_out9,
                             void* stream) {
    if (int rc = check_common(device, d_sorted_cols, d_out9, n_regulons, n_cells)) return rc;
    if (!d_split || max_iter < 0) return fail(AUCELL_ERR_INVALID, "binarize: bad split / max_iter");
    DeviceGuard dg(device);
    if (!dg.ok) return fail(AUCELL_ERR_CUDA, "cudaSetDevice failed");
    binarize_gmm2_kernel<<<n_regulons, kBT, 0, (cudaStream_t)stream>>>(d_sorted_cols, n_cells, d_split, max_iter, tol,
                                                                       reg_covar, d_out9);
    launch_counter().fetch_add(1);
    CK(cudaGetLastError());
    return AUCELL_OK;
}

int aucell_binarize_trough_f64(int device, const double* d_cols, int32_t n_regulons, int64_t n_cells, const double* d_lo,
                               const double* d_hi, const double* d_kde_var, double xatol, int32_t maxiter,
                               double* d_out_x, void* stream) {
    if (int rc = check_common(device, d_cols, d_out_x, n_regulons, n_cells)) return rc;
    if (!d_lo || !d_hi || !d_kde_var) return fail(AUCELL_ERR_INVALID, "NULL pointer");
    DeviceGuard dg(device);
    if (!dg.ok) return fail(AUCELL_ERR_CUDA, "cudaSetDevice failed");
    binarize_trough_kernel<<<n_regulons, kBT, 0, (cudaStream_t)stream>>>(d_cols, n_cells, d_lo, d_hi, d_kde_var, xatol,
                                                                         maxiter, d_out_x);
    launch_counter().fetch_add(1);
    CK(cudaGetLastError());
    return AUCELL_OK;
}

}  // extern "C"
