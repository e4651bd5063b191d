/*
 * oracle.c -- CPU restatement of damavand's `multithreading` apply method, measure,
 * sampler and expectation-value extraction.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / `--impl reference` legs may load it.  The product
 * (damavand_b200/) never imports, links or executes anything in oracle/.
 *
 * Parity status: the reference (Rust + rsmpi + PyO3) cannot be compiled in this image
 * (no cargo/rustc/MPI), so this file is a line-by-line restatement.  It is pinned against
 * the ONE known-answer vector the reference's own tests hold
 * (src/qubit_backend/circuit.rs:803-837: H(0),H(1),CNOT(0,1) -> [.5,.5,.5,.5]) and is
 * cross-checked against an independent restatement of the reference's `brute_force`
 * method (oracle/oracle.py).  Beyond that vector: PARITY UNPINNED by reference tests.
 *
 * Build: gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC  (no FMA contraction, like rustc).
 * All amplitudes are interleaved complex128: amp[2*i] = re, amp[2*i+1] = im.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { double re, im; } c64;

/* num::Complex<f64> Mul: (a.re*b.re - a.im*b.im, a.re*b.im + a.im*b.re) */
static inline c64 cmul(c64 a, c64 b) {
    c64 r; r.re = a.re * b.re - a.im * b.im; r.im = a.re * b.im + a.im * b.re; return r;
}
static inline c64 cadd(c64 a, c64 b) { c64 r; r.re = a.re + b.re; r.im = a.im + b.im; return r; }

/* Circuit::compute_partner_rank, src/qubit_backend/circuit.rs:781-795 */
uint64_t orc_compute_partner_rank(uint64_t current, uint64_t per_node, uint64_t gap) {
    uint64_t before = current * per_node;
    if (before % (2 * gap) < gap) return current + gap / per_node;
    return current - gap / per_node;
}

/*
 * apply_multithreading_local, src/qubit_backend/circuit_multithreading.rs:9-54.
 *   m = [m00.re,m00.im, m01.re,m01.im, m10.re,m10.im, m11.re,m11.im]  (row-major 2x2)
 *   control < 0 means "no control qubit".
 *   scratch must hold n_amps complex values: it is the `partner_amplitudes = clone()` (:31).
 */
void orc_apply_gate(double *amp_, double *scratch_, uint64_t n_amps, const double *m,
                    int control, int target) {
    c64 *amp = (c64 *)amp_;
    c64 *partner = (c64 *)scratch_;
    const c64 m00 = {m[0], m[1]}, m01 = {m[2], m[3]}, m10 = {m[4], m[5]}, m11 = {m[6], m[7]};
    const uint64_t gap = 1ull << target;
    memcpy(partner, amp, n_amps * sizeof(c64));          /* :31 clone */
    #pragma omp parallel for schedule(static)
    for (uint64_t i = 0; i < n_amps; ++i) {               /* :33 par_for_each */
        uint64_t apply = control >= 0 ? ((i >> control) & 1ull) : 1ull;   /* :35-39 */
        uint64_t p = orc_compute_partner_rank(i, 1, gap);                  /* :40-41 */
        if (apply == 1ull) {
            c64 upper = partner[p];                                        /* :44 */
            if (p > i) amp[i] = cadd(cmul(amp[i], m00), cmul(upper, m01)); /* :47 */
            else       amp[i] = cadd(cmul(upper, m10), cmul(amp[i], m11)); /* :49 */
        }
    }
}

/* Circuit::measure CPU branch, src/qubit_backend/circuit.rs:579-581 (norm_sqr = re*re+im*im) */
void orc_measure(const double *amp_, uint64_t n_amps, double *probs) {
    const c64 *amp = (const c64 *)amp_;
    #pragma omp parallel for schedule(static)
    for (uint64_t i = 0; i < n_amps; ++i) probs[i] = amp[i].re * amp[i].re + amp[i].im * amp[i].im;
}

/*
 * utils::sample_from_discrete_distribution + sample_from_discrete_cumulative,
 * src/utils.rs:258-277, with thread_rng().gen::<f64>() replaced by the injected u[s].
 * faithful != 0 : rebuild the cumulative array for every shot and scan it linearly, exactly
 *                 as the reference does (O(shots*N)); used for the timed CPU baseline.
 * faithful == 0 : build the (identical, deterministic) cumulative once, same linear-scan rule
 *                 evaluated by lower-bound search (cum is non-decreasing, so the first index with
 *                 xsi <= cum[index] is the same).
 * Result = index-1 (utils.rs:264).  xsi==0 hits index 0 and the reference underflows usize;
 * we define that case as 0.
 */
void orc_sample_sequential(const double *probs, uint64_t n, const double *u, uint64_t shots,
                           uint64_t *out, int faithful) {
    double *cum = (double *)malloc((n + 1) * sizeof(double));
    if (!faithful) {
        cum[0] = 0.0;
        for (uint64_t j = 0; j < n; ++j) cum[j + 1] = cum[j] + probs[j];   /* utils.rs:270-274 */
    }
    for (uint64_t s = 0; s < shots; ++s) {
        if (faithful) {
            cum[0] = 0.0;
            for (uint64_t j = 0; j < n; ++j) cum[j + 1] = cum[j] + probs[j];
        }
        double xsi = u[s] * cum[n];                                        /* utils.rs:260 */
        uint64_t index;
        if (faithful) {
            for (index = 0; index <= n; ++index) if (xsi <= cum[index]) break;   /* :262-266 */
        } else {
            uint64_t lo = 0, hi = n;          /* smallest index with xsi <= cum[index] */
            while (lo < hi) { uint64_t mid = (lo + hi) >> 1; if (xsi <= cum[mid]) hi = mid; else lo = mid + 1; }
            index = lo;
        }
        out[s] = index == 0 ? 0 : index - 1;
    }
    free(cum);
}

/*
 * Pairwise-tree sampler: the SPECIFIED summation order of the CUDA sampler
 * (documented deviation from the reference's strictly sequential cumulative sum, which
 * cannot be evaluated in parallel bit-for-bit).  tree level 0 = probs; level l+1[j] =
 * level l[2j] + level l[2j+1].  total = root.  Descent for xsi = u*total: at a node with
 * running base b, c = b + left_child; xsi <= c -> go left, else b = c and go right.
 * n must be a power of two.
 */
void orc_sample_tree(const double *probs, uint64_t n, const double *u, uint64_t shots,
                     uint64_t *out) {
    /* levels stored back to back: level 0 (n), level 1 (n/2), ... root (1) */
    int levels = 0; while ((1ull << levels) < n) ++levels;
    double **lv = (double **)malloc((levels + 1) * sizeof(double *));
    lv[0] = (double *)probs;
    for (int l = 1; l <= levels; ++l) {
        uint64_t cnt = n >> l;
        lv[l] = (double *)malloc(cnt * sizeof(double));
        const double *src = lv[l - 1];
        double *dst = lv[l];
        #pragma omp parallel for schedule(static)
        for (uint64_t j = 0; j < cnt; ++j) dst[j] = src[2 * j] + src[2 * j + 1];
    }
    const double total = lv[levels][0];
    for (uint64_t s = 0; s < shots; ++s) {
        double xsi = u[s] * total;
        double base = 0.0;
        uint64_t j = 0;
        for (int l = levels; l >= 1; --l) {
            double c = base + lv[l - 1][2 * j];
            if (xsi <= c) j = 2 * j; else { base = c; j = 2 * j + 1; }
        }
        out[s] = j;
    }
    for (int l = 1; l <= levels; ++l) free(lv[l]);
    free(lv);
}

double orc_tree_total(const double *probs, uint64_t n) {
    if (n == 1) return probs[0];
    uint64_t cnt = n >> 1;
    double *buf = (double *)malloc(cnt * sizeof(double));
    for (uint64_t j = 0; j < cnt; ++j) buf[j] = probs[2 * j] + probs[2 * j + 1];
    while (cnt > 1) { cnt >>= 1; for (uint64_t j = 0; j < cnt; ++j) buf[j] = buf[2 * j] + buf[2 * j + 1]; }
    double t = buf[0]; free(buf); return t;
}

/* Circuit::extract_expectation_values, src/qubit_backend/circuit.rs:494-513.
 * out is [shots][n_obs] row-major. */
void orc_extract_expectation_values(const uint64_t *samples, uint64_t shots, const int *qubits,
                                    int n_obs, double *out) {
    for (uint64_t s = 0; s < shots; ++s)
        for (int o = 0; o < n_obs; ++o)
            out[s * (uint64_t)n_obs + o] = ((samples[s] >> qubits[o]) & 1ull) > 0 ? -1.0 : 1.0;
}
