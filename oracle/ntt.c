/*
 * oracle/ntt.c -- radix-2 NTT / iNTT / coset LDE over Goldilocks and the column commitment.
 * TEST INFRASTRUCTURE ONLY (see oracle/gl.h header).
 *
 * Restates plonky2_field 0.2.0 fft.rs (`fft` = bit-reverse then in-place DIT butterflies with a
 * per-stage root table; `ifft` = fft, reverse entries 1..n, scale by 1/n) and plonky2
 * fri/oracle.rs PolynomialBatch::{from_values, lde_values}: values -> ifft -> zero-pad by 2^r ->
 * coset fft (shift 7) -> transpose -> bit-reverse rows -> Merkle.  Source not under
 * /root/reference (Cargo.lock:2957-2982); reached via `circuit.prove()` [REF circuits/skip.rs:214].
 * Pinned by the NTT8 known answer in SURVEY.md App. C and an O(n^2) DFT cross-check.
 */
#include "oracle.h"
#include <stdlib.h>
#include <string.h>

static void bitrev_permute(gl_t *a, size_t n) {
    unsigned lg = tmx_log2(n);
    for (size_t i = 0; i < n; i++) {
        size_t j = tmx_bitrev(i, lg);
        if (i < j) {
            gl_t t = a[i];
            a[i] = a[j];
            a[j] = t;
        }
    }
}

/* twiddle cache: tw[k] holds w_{2^k}^i for i < 2^(k-1) */
#define MAX_LG 28
static gl_t *tw_cache[MAX_LG + 1];

static const gl_t *twiddles(unsigned k) {
    gl_t *t;
#pragma omp critical(tw_cache_lock)
    {
        t = tw_cache[k];
        if (!t) {
            size_t half = (size_t)1 << (k - 1);
            t = (gl_t *)malloc(half * sizeof(gl_t));
            gl_t w = gl_root_of_unity(k), cur = 1;
            for (size_t i = 0; i < half; i++) {
                t[i] = cur;
                cur = gl_mul(cur, w);
            }
            tw_cache[k] = t;
        }
    }
    return t;
}

void ntt_forward(gl_t *a, size_t n) {
    if (n <= 1) return;
    unsigned lg = tmx_log2(n);
    bitrev_permute(a, n);
    const gl_t *tw = twiddles(lg); /* w_n^i, i < n/2 */
    for (unsigned s = 1; s <= lg; s++) {
        size_t m = (size_t)1 << s, half = m >> 1, stride = n / m;
        for (size_t k = 0; k < n; k += m)
            for (size_t j = 0; j < half; j++) {
                gl_t u = a[k + j];
                gl_t v = gl_mul(a[k + j + half], tw[j * stride]);
                a[k + j] = gl_add(u, v);
                a[k + j + half] = gl_sub(u, v);
            }
    }
}

void ntt_inverse(gl_t *a, size_t n) {
    if (n <= 1) return;
    ntt_forward(a, n);
    for (size_t i = 1, j = n - 1; i < j; i++, j--) {
        gl_t t = a[i];
        a[i] = a[j];
        a[j] = t;
    }
    gl_t ninv = gl_inv((gl_t)n);
    for (size_t i = 0; i < n; i++) a[i] = gl_mul(a[i], ninv);
}

void ntt_naive_dft(const gl_t *in, gl_t *out, size_t n) {
    gl_t w = gl_root_of_unity(tmx_log2(n));
    for (size_t k = 0; k < n; k++) {
        gl_t wk = gl_pow(w, k), cur = 1, acc = 0;
        for (size_t j = 0; j < n; j++) {
            acc = gl_add(acc, gl_mul(in[j], cur));
            cur = gl_mul(cur, wk);
        }
        out[k] = acc;
    }
}

void ntt_coset_forward(gl_t *a, size_t n, gl_t shift) {
    gl_t cur = 1;
    for (size_t i = 0; i < n; i++) {
        a[i] = gl_mul(a[i], cur);
        cur = gl_mul(cur, shift);
    }
    ntt_forward(a, n);
}

void ntt_coset_inverse(gl_t *a, size_t n, gl_t shift) {
    ntt_inverse(a, n);
    gl_t si = gl_inv(shift), cur = 1;
    for (size_t i = 0; i < n; i++) {
        a[i] = gl_mul(a[i], cur);
        cur = gl_mul(cur, si);
    }
}

void ntt_lde_batch(const gl_t *values, size_t n_cols, size_t n, unsigned rate_bits, gl_t *out, gl_t *coeffs_out) {
    size_t m = n << rate_bits;
    unsigned lgm = tmx_log2(m);
    (void)twiddles(tmx_log2(n) ? tmx_log2(n) : 1);
    (void)twiddles(lgm ? lgm : 1);
#pragma omp parallel
    {
        gl_t *buf = (gl_t *)malloc(m * sizeof(gl_t));
#pragma omp for schedule(dynamic, 4)
        for (size_t c = 0; c < n_cols; c++) {
            memcpy(buf, values + c * n, n * sizeof(gl_t));
            ntt_inverse(buf, n);
            if (coeffs_out) memcpy(coeffs_out + c * n, buf, n * sizeof(gl_t));
            memset(buf + n, 0, (m - n) * sizeof(gl_t));
            ntt_coset_forward(buf, m, GL_GENERATOR);
            gl_t *o = out + c * m;
            for (size_t j = 0; j < m; j++) o[j] = buf[tmx_bitrev(j, lgm)];
        }
        free(buf);
    }
}

void commit_columns(merkle_tree_t *t, const gl_t *cols, size_t n_cols, size_t n_rows, unsigned cap_height) {
    /* transpose to rows (the leaves), then hash */
    gl_t *rows = (gl_t *)malloc(n_cols * n_rows * sizeof(gl_t));
#pragma omp parallel for schedule(static)
    for (size_t j = 0; j < n_rows; j++)
        for (size_t c = 0; c < n_cols; c++) rows[j * n_cols + c] = cols[c * n_rows + j];
    merkle_build(t, rows, n_rows, n_cols, cap_height);
    free(rows);
}
