// ref_shim.cpp — C-ABI binding of the REFERENCE's own C/C++ sources (TEST INFRASTRUCTURE ONLY).
//
// Compiled by oracle/Makefile together with /root/reference/src/include/{bpgd.cpp,mod2sparse.c,
// mod2sparse_extra.cpp} (where they lie; nothing is copied) into oracle/_ref/libswd_ref.so.
// It lets the tests and the CPU baseline call the real BPGD_main_thread::do_work (threads and all),
// index_sort, mod2sparse_decomp_osd and LU_forward_backward_solve.  The only logic restated here is
// what lives in the reference's Cython layer and therefore cannot be compiled without Cython:
// the unmasked pre-BP loop (bp_guessing_decoder.pyx:48-139) and the decode() glue (pyx:221-251),
// both written against the reference's mod2sparse container.
#include <cstdio>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "bpgd.hpp"
#include "mod2sparse_extra.hpp"

struct RefGdg {
    int m, n, new_n, max_iter;
    double alpha;
    mod2sparse *H;
    std::vector<double> llr, llr_sum;
    std::vector<double> hist;            // [n][4]
    std::vector<char> synd, dec, dsynd;
    std::vector<int> cols;
    BPGD_main_thread *mt;
};

extern "C" {

void *ref_gdg_create(int m, int n, const int *cp, const int *cr, const double *llr, int max_iter, double ms_scaling_factor,
                     int max_iter_per_step, int max_step, int max_tree_depth, int max_side_depth, int max_tree_branch_step,
                     int max_side_branch_step, double gdg_factor, int new_n, int low_error_mode) {
    RefGdg *r = new RefGdg();
    r->m = m; r->n = n; r->max_iter = max_iter; r->alpha = ms_scaling_factor;
    r->new_n = new_n <= 0 ? (n < 2 * m ? n : 2 * m) : (new_n < n ? new_n : n);
    r->H = mod2sparse_allocate(m, n);
    for (int c = 0; c < n; c++) for (int e = cp[c]; e < cp[c + 1]; e++) mod2sparse_insert(r->H, cr[e], c);
    r->llr.assign(llr, llr + n); r->llr_sum.assign(n, 0.0); r->hist.assign((size_t)4 * n, 0.0);
    r->synd.assign(m, 0); r->dsynd.assign(m, 0); r->dec.assign(n, 0); r->cols.assign(n, 0);
    r->mt = new BPGD_main_thread(m, r->new_n, max_iter_per_step, max_step, max_tree_depth, max_side_depth, max_tree_branch_step,
                                 max_side_branch_step, low_error_mode, gdg_factor);
    return r;
}

void ref_gdg_destroy(void *h) {
    RefGdg *r = (RefGdg *)h;
    if (!r) return;
    delete r->mt;
    mod2sparse_free(r->H);
    delete r;
}

// pre-BP exactly as bp_history_decoder.bp_decode_llr (pyx:48-139), on the reference's linked-list matrix
static int ref_pre_bp(RefGdg *r) {
    mod2sparse *H = r->H;
    mod2entry *e;
    for (int vn = 0; vn < r->n; vn++)
        for (e = mod2sparse_first_in_col(H, vn); !mod2sparse_at_end(e); e = mod2sparse_next_in_col(e)) e->bit_to_check = r->llr[vn];
    for (int it = 0; it < r->max_iter; it++) {
        for (int cn = 0; cn < r->m; cn++) {
            double temp = 1e308; int sgn = r->synd[cn] == 1 ? 1 : 0;
            for (e = mod2sparse_first_in_row(H, cn); !mod2sparse_at_end(e); e = mod2sparse_next_in_row(e)) {
                e->check_to_bit = temp; e->sgn = sgn;
                if (e->bit_to_check > 50.0) e->bit_to_check = 50.0; else if (e->bit_to_check < -50.0) e->bit_to_check = -50.0;
                if (std::fabs(e->bit_to_check) < temp) temp = std::fabs(e->bit_to_check);
                if (e->bit_to_check <= 0) sgn = 1 - sgn;
            }
            temp = 1e308; sgn = 0;
            for (e = mod2sparse_last_in_row(H, cn); !mod2sparse_at_end(e); e = mod2sparse_prev_in_row(e)) {
                if (temp < e->check_to_bit) e->check_to_bit = temp;
                e->sgn += sgn;
                e->check_to_bit *= ((e->sgn % 2 == 0) ? 1.0 : -1.0) * r->alpha;
                if (std::fabs(e->bit_to_check) < temp) temp = std::fabs(e->bit_to_check);
                if (e->bit_to_check <= 0) sgn = 1 - sgn;
            }
        }
        for (int vn = 0; vn < r->n; vn++) {
            double temp = r->llr[vn];
            for (e = mod2sparse_first_in_col(H, vn); !mod2sparse_at_end(e); e = mod2sparse_next_in_col(e)) { e->bit_to_check = temp; temp += e->check_to_bit; }
            r->hist[(size_t)4 * vn + it % 4] = temp;
            r->dec[vn] = temp <= 0 ? 1 : 0;
            temp = 0.0;
            for (e = mod2sparse_last_in_col(H, vn); !mod2sparse_at_end(e); e = mod2sparse_prev_in_col(e)) { e->bit_to_check += temp; temp += e->check_to_bit; }
        }
        mod2sparse_mulvec(H, r->dec.data(), r->dsynd.data());
        if (memcmp(r->dsynd.data(), r->synd.data(), r->m) == 0) return 1;
    }
    return 0;
}

// bpgdg_decoder.decode with multi_thread=True (pyx:221-251): pre-BP, index_sort, the REAL do_work
int ref_gdg_decode(void *h, const signed char *synd, signed char *dec, double *min_pm) {
    RefGdg *r = (RefGdg *)h;
    memcpy(r->synd.data(), synd, r->m);
    int conv = ref_pre_bp(r);
    if (min_pm) *min_pm = 10000.0;
    if (!conv) {
        for (int vn = 0; vn < r->n; vn++) {
            const double *q = &r->hist[(size_t)4 * vn];
            r->llr_sum[vn] = q[0] + q[1] + q[2] + q[3];
        }
        index_sort(r->llr_sum.data(), r->cols.data(), r->n);
        r->mt->do_work(r->H, r->cols.data(), r->llr.data(), r->synd.data());
        conv = r->mt->min_pm < 9999.0;
        for (int vn = 0; vn < r->new_n; vn++) r->dec[r->cols[vn]] = r->mt->min_pm_error[vn];
        for (int vn = r->new_n; vn < r->n; vn++) r->dec[r->cols[vn]] = 0;
        if (min_pm) *min_pm = r->mt->min_pm;
    }
    memcpy(dec, r->dec.data(), r->n);
    return conv;
}

void ref_gdg_decode_batch(void *h, const signed char *synd, long long B, signed char *dec, signed char *conv) {
    RefGdg *r = (RefGdg *)h;
    for (long long b = 0; b < B; b++) conv[b] = (signed char)ref_gdg_decode(h, synd + b * r->m, dec + b * r->n, nullptr);
}

// index_sort (bpgd.cpp:384-389)
void ref_index_sort(double *v, int *cols, int n) { index_sort(v, cols, n); }

// mod2sparse_rank (mod2sparse_extra.cpp:32-76)
int ref_rank(int m, int n, const int *cp, const int *cr) {
    mod2sparse *H = mod2sparse_allocate(m, n);
    for (int c = 0; c < n; c++) for (int e = cp[c]; e < cp[c + 1]; e++) mod2sparse_insert(H, cr[e], c);
    int rk = mod2sparse_rank(H);
    mod2sparse_free(H);
    return rk;
}

// OSD-0 through the reference's LU (osd_window.pyx:215-229): cols[] in: column order, out: pivots first; x out
int ref_osd0(int m, int n, const int *cp, const int *cr, int rank, int *cols, const signed char *synd, signed char *x) {
    mod2sparse *H = mod2sparse_allocate(m, n);
    for (int c = 0; c < n; c++) for (int e = cp[c]; e < cp[c + 1]; e++) mod2sparse_insert(H, cr[e], c);
    mod2sparse *L = mod2sparse_allocate(m, rank), *U = mod2sparse_allocate(rank, n);
    std::vector<int> rows(m);
    std::vector<char> z(synd, synd + m), xx(n, 0);
    int nnf = mod2sparse_decomp_osd(H, rank, L, U, rows.data(), cols);
    LU_forward_backward_solve(L, U, rows.data(), cols, z.data(), xx.data());
    memcpy(x, xx.data(), n);
    mod2sparse_free(L); mod2sparse_free(U); mod2sparse_free(H);
    return nnf;
}

}  // extern "C"
