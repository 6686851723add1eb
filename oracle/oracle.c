/*
 * oracle.c — CPU restatement of the reference algorithm for the block-sparse contraction path.
 *
 * TEST INFRASTRUCTURE ONLY. Nothing in the product (itensor_b200/, include/) may import, link or call
 * this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * use it, and only as the checker. It is deliberately the plain definition (index loops), written
 * independently of the CUDA planner (itensor_b200/csrc/plan.cc).
 *
 * Parity pin: checked against the UNMODIFIED reference built into oracle/_ref/libitref.so
 * (oracle/Makefile, oracle/ref_harness.cc) by tests/test_oracle_vs_reference.py, and against the
 * committed golden vectors in tests/golden/ (generated from the reference by
 * tests/golden/make_golden.py).
 *
 * What each function follows (paths relative to the ITensor v3 tree):
 *   orc_contract_structure  computeLabels            itensor/tensor/contract.h:155-202
 *                           contractIS(sort=false)   itensor/indexset_impl.h:77-129
 *                           getContractedOffsets     itensor/itdata/qutil.h:93-242
 *                           Block operator<          itensor/itdata/qdense.cc:60-65
 *   orc_contract_values     doTask(Contract,QDense,QDense) + do_contract lambda
 *                                                    itensor/itdata/qdense.cc:671-747
 *                           (per pair: C_block = or += A_block*B_block, beta 0 then 1, :727-731)
 *                           Dense case = one block   itensor/itdata/dense.cc:262-330
 *   orc_permute             permuteQDense / add()    itensor/itdata/qdense.cc:515-549,847-881
 *                           transform()              itensor/tensor/ten_impl.h:107-160
 *   orc_nrm2                dnrm2 (reference BLAS scaled sum of squares) via
 *                           doTask(NormNoScale)      itensor/itdata/qdense.cc:409-417
 *   orc_flux_blocks         getBlockOffsets          itensor/itdata/qdense.cc:133-173
 *                           QNum::set (mod rule)     itensor/qn.cc:10-29
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_MAXR 16

typedef struct {
    int32_t order, dtype; /* dtype 0 real, 1 complex (interleaved) */
    const int32_t* nsect;
    const int64_t* sect;
    int64_t nblocks;
    const int32_t* blocks;
    const int64_t* offsets;
    int64_t nelems;
} orc_desc;

static int64_t sect_size(const orc_desc* t, int j, int s) {
    int64_t st = 0;
    for (int i = 0; i < j; ++i) st += t->nsect[i];
    return t->sect[st + s];
}

/* reference Block ordering: compare from the LAST coordinate backwards */
static int block_cmp(const int32_t* a, const int32_t* b, int r) {
    for (int j = r - 1; j >= 0; --j) {
        if (a[j] < b[j]) return -1;
        if (a[j] > b[j]) return 1;
    }
    return 0;
}

static int g_rC; /* qsort context (single-threaded test helper) */
static int cmp_rows(const void* x, const void* y) { return block_cmp((const int32_t*)x, (const int32_t*)y, g_rC); }

/* computeLabels: AtoB[i] = j of the first B index with the same label (each B index used once) */
static void match_labels(int rA, const int32_t* labA, int rB, const int32_t* labB, int* AtoB, int* BtoA) {
    for (int i = 0; i < rA; ++i) AtoB[i] = -1;
    for (int j = 0; j < rB; ++j) BtoA[j] = -1;
    for (int i = 0; i < rA; ++i)
        for (int j = 0; j < rB; ++j)
            if (labA[i] == labB[j] && BtoA[j] < 0) { AtoB[i] = j; BtoA[j] = i; break; }
}

/*
 * Structure of C = A*B. Caller provides capacity; returns 0 on success, -1 if a capacity is too small
 * (the needed counts are still written). triples are (iA, iB, iC) in the reference's enumeration order
 * (A blocks outer, B blocks inner).
 */
int orc_contract_structure(const orc_desc* A, const int32_t* labA, const orc_desc* B, const int32_t* labB,
                           int32_t* c_order, int32_t* c_labels, int32_t* c_nsect, int64_t* c_sect,
                           int64_t* npairs, int64_t* triples, int64_t cap_pairs, int64_t* c_nblocks,
                           int32_t* c_blocks, int64_t* c_offsets, int64_t cap_blocks, int64_t* c_nelems) {
    const int rA = A->order, rB = B->order;
    int AtoB[ORC_MAXR], BtoA[ORC_MAXR], AtoC[ORC_MAXR], BtoC[ORC_MAXR];
    match_labels(rA, labA, rB, labB, AtoB, BtoA);
    /* contractIS: uncontracted A indices in A order, then uncontracted B indices in B order */
    int rC = 0;
    int64_t ns = 0;
    for (int i = 0; i < rA; ++i) {
        AtoC[i] = -1;
        if (AtoB[i] < 0) {
            AtoC[i] = rC;
            c_labels[rC] = labA[i];
            c_nsect[rC] = A->nsect[i];
            for (int s = 0; s < A->nsect[i]; ++s) c_sect[ns++] = sect_size(A, i, s);
            ++rC;
        }
    }
    for (int j = 0; j < rB; ++j) {
        BtoC[j] = -1;
        if (BtoA[j] < 0) {
            BtoC[j] = rC;
            c_labels[rC] = labB[j];
            c_nsect[rC] = B->nsect[j];
            for (int s = 0; s < B->nsect[j]; ++s) c_sect[ns++] = sect_size(B, j, s);
            ++rC;
        }
    }
    *c_order = rC;
    /* getContractedOffsets: O(nA*nB) scan, match on contracted block coordinates */
    int64_t np = 0;
    const int rowC = rC > 0 ? rC : 1;
    int32_t* cbl = (int32_t*)malloc(sizeof(int32_t) * (size_t)rowC * (size_t)(cap_pairs > 0 ? cap_pairs : 1));
    for (int64_t a = 0; a < A->nblocks; ++a) {
        const int32_t* ab = A->blocks + a * rA;
        for (int64_t b = 0; b < B->nblocks; ++b) {
            const int32_t* bb = B->blocks + b * rB;
            int ok = 1;
            for (int i = 0; i < rA && ok; ++i)
                if (AtoB[i] >= 0 && ab[i] != bb[AtoB[i]]) ok = 0;
            if (!ok) continue;
            if (np < cap_pairs) {
                triples[3 * np] = a;
                triples[3 * np + 1] = b;
                int32_t* row = cbl + np * rowC;
                row[0] = 0;
                for (int i = 0; i < rA; ++i) if (AtoC[i] >= 0) row[AtoC[i]] = ab[i];
                for (int j = 0; j < rB; ++j) if (BtoC[j] >= 0) row[BtoC[j]] = bb[j];
            }
            ++np;
        }
    }
    *npairs = np;
    if (np > cap_pairs) { free(cbl); return -1; }
    /* sort + unique the C block labels, then prefix-sum block sizes into offsets */
    int32_t* sorted = (int32_t*)malloc(sizeof(int32_t) * (size_t)rowC * (size_t)(np > 0 ? np : 1));
    memcpy(sorted, cbl, sizeof(int32_t) * (size_t)rowC * (size_t)np);
    g_rC = rC;
    qsort(sorted, (size_t)np, sizeof(int32_t) * (size_t)rowC, cmp_rows);
    int64_t nb = 0;
    for (int64_t p = 0; p < np; ++p) {
        if (p > 0 && block_cmp(sorted + p * rowC, sorted + (p - 1) * rowC, rC) == 0) continue;
        if (nb < cap_blocks) memcpy(sorted + nb * rowC, sorted + p * rowC, sizeof(int32_t) * (size_t)rowC);
        ++nb;
    }
    *c_nblocks = nb;
    if (nb > cap_blocks) { free(cbl); free(sorted); return -1; }
    int64_t off = 0;
    {
        int64_t cst[ORC_MAXR + 1];
        cst[0] = 0;
        for (int j = 0; j < rC; ++j) cst[j + 1] = cst[j] + c_nsect[j];
        for (int64_t c = 0; c < nb; ++c) {
            int64_t sz = 1;
            for (int j = 0; j < rC; ++j) {
                c_blocks[c * rC + j] = sorted[c * rowC + j];
                sz *= c_sect[cst[j] + sorted[c * rowC + j]];
            }
            c_offsets[c] = off;
            off += sz;
        }
    }
    *c_nelems = off;
    /* iC of each pair: binary search in the sorted unique list (offsetOf, qdense.cc:214-221) */
    for (int64_t p = 0; p < np; ++p) {
        int64_t lo = 0, hi = nb - 1, pos = -1;
        while (lo <= hi) {
            int64_t mid = (lo + hi) / 2;
            int c = block_cmp(sorted + mid * rowC, cbl + p * rowC, rC);
            if (c == 0) { pos = mid; break; }
            if (c < 0) lo = mid + 1; else hi = mid - 1;
        }
        triples[3 * p + 2] = pos;
    }
    free(cbl);
    free(sorted);
    return 0;
}

/* multi-index odometer helpers: offsets of every element of an index subset */
static int64_t fill_offsets(int n, const int64_t* ext, const int64_t* str, int64_t** out) {
    int64_t tot = 1;
    for (int d = 0; d < n; ++d) tot *= ext[d];
    int64_t* o = (int64_t*)malloc(sizeof(int64_t) * (size_t)(tot > 0 ? tot : 1));
    int64_t idx[ORC_MAXR] = {0};
    for (int64_t e = 0; e < tot; ++e) {
        int64_t v = 0;
        for (int d = 0; d < n; ++d) v += idx[d] * str[d];
        o[e] = v;
        for (int d = 0; d < n; ++d) {
            if (++idx[d] < ext[d]) break;
            idx[d] = 0;
        }
    }
    *out = o;
    return tot;
}

/* C (pre-allocated, c structure from orc_contract_structure) = A*B, pair by pair in the given order */
int orc_contract_values(const orc_desc* A, const int32_t* labA, const double* Ad, const orc_desc* B,
                        const int32_t* labB, const double* Bd, const orc_desc* C, const int64_t* triples,
                        int64_t npairs, double* Cd) {
    const int rA = A->order, rB = B->order;
    int AtoB[ORC_MAXR], BtoA[ORC_MAXR];
    match_labels(rA, labA, rB, labB, AtoB, BtoA);
    const int ca = A->dtype == 1, cb = B->dtype == 1, cc = ca || cb;
    char* touched = (char*)calloc((size_t)(C->nblocks > 0 ? C->nblocks : 1), 1);
    for (int64_t p = 0; p < npairs; ++p) {
        const int64_t ia = triples[3 * p], ib = triples[3 * p + 1], ic = triples[3 * p + 2];
        const int32_t* ab = A->blocks + ia * rA;
        const int32_t* bb = B->blocks + ib * rB;
        int64_t eA[ORC_MAXR], sA[ORC_MAXR], eB[ORC_MAXR], sB[ORC_MAXR];
        int64_t s = 1;
        for (int i = 0; i < rA; ++i) { eA[i] = sect_size(A, i, ab[i]); sA[i] = s; s *= eA[i]; }
        s = 1;
        for (int j = 0; j < rB; ++j) { eB[j] = sect_size(B, j, bb[j]); sB[j] = s; s *= eB[j]; }
        /* M: uncontracted A (A order); N: uncontracted B (B order); K: contracted (A order) */
        int64_t me[ORC_MAXR], ms[ORC_MAXR], ne[ORC_MAXR], nst[ORC_MAXR], ke[ORC_MAXR], kas[ORC_MAXR], kbs[ORC_MAXR];
        int nm = 0, nn = 0, nk = 0;
        for (int i = 0; i < rA; ++i) {
            if (AtoB[i] < 0) { me[nm] = eA[i]; ms[nm] = sA[i]; ++nm; }
            else { ke[nk] = eA[i]; kas[nk] = sA[i]; kbs[nk] = sB[AtoB[i]]; ++nk; }
        }
        for (int j = 0; j < rB; ++j)
            if (BtoA[j] < 0) { ne[nn] = eB[j]; nst[nn] = sB[j]; ++nn; }
        int64_t *om, *on, *oka, *okb;
        const int64_t M = fill_offsets(nm, me, ms, &om);
        const int64_t N = fill_offsets(nn, ne, nst, &on);
        const int64_t K = fill_offsets(nk, ke, kas, &oka);
        fill_offsets(nk, ke, kbs, &okb);
        const double* a = Ad + A->offsets[ia] * (ca ? 2 : 1);
        const double* b = Bd + B->offsets[ib] * (cb ? 2 : 1);
        double* c = Cd + C->offsets[ic] * (cc ? 2 : 1);
        const int first = !touched[ic]; /* beta = 0 on first touch, 1 afterwards */
        touched[ic] = 1;
        for (int64_t n = 0; n < N; ++n) {
            for (int64_t m = 0; m < M; ++m) {
                double sr = 0.0, si = 0.0;
                for (int64_t k = 0; k < K; ++k) {
                    const int64_t oa = om[m] + oka[k], ob = okb[k] + on[n];
                    const double ar = ca ? a[2 * oa] : a[oa], ai = ca ? a[2 * oa + 1] : 0.0;
                    const double br = cb ? b[2 * ob] : b[ob], bi = cb ? b[2 * ob + 1] : 0.0;
                    sr += ar * br - ai * bi;
                    si += ar * bi + ai * br;
                }
                const int64_t oc = m + M * n; /* C index order = [M indices][N indices], column-major */
                if (cc) {
                    if (first) { c[2 * oc] = sr; c[2 * oc + 1] = si; }
                    else { c[2 * oc] += sr; c[2 * oc + 1] += si; }
                } else {
                    if (first) c[oc] = sr; else c[oc] += sr;
                }
            }
        }
        free(om); free(on); free(oka); free(okb);
    }
    free(touched);
    return 0;
}

/*
 * dst (=|+=) alpha * permute(src). perm[i] = position in dst of src index i (Permutation::dest).
 * Returns -1 if a src block has no image in dst. Blocks of dst without a source are zeroed when
 * !accumulate (permuteQDense allocates zero-initialised storage for every flux-allowed block).
 */
int orc_permute(const orc_desc* S, const double* Sd, const orc_desc* D, double* Dd, const int32_t* perm,
                double alpha_re, double alpha_im, int accumulate) {
    const int r = S->order;
    const int cs = S->dtype == 1, cd = D->dtype == 1;
    if (!accumulate) memset(Dd, 0, sizeof(double) * (size_t)D->nelems * (cd ? 2 : 1));
    for (int64_t b = 0; b < S->nblocks; ++b) {
        const int32_t* sb = S->blocks + b * r;
        int32_t db[ORC_MAXR];
        for (int i = 0; i < r; ++i) db[perm[i]] = sb[i];
        int64_t pos = -1;
        for (int64_t q = 0; q < D->nblocks; ++q)
            if (r == 0 || block_cmp(D->blocks + q * r, db, r) == 0) { pos = q; break; }
        if (pos < 0) return -1;
        int64_t eS[ORC_MAXR], dstr[ORC_MAXR], dS[ORC_MAXR];
        int64_t tot = 1;
        for (int i = 0; i < r; ++i) { eS[i] = sect_size(S, i, sb[i]); tot *= eS[i]; }
        int64_t s = 1;
        for (int j = 0; j < r; ++j) { dS[j] = s; s *= sect_size(D, j, db[j]); }
        for (int i = 0; i < r; ++i) dstr[i] = dS[perm[i]];
        int64_t idx[ORC_MAXR] = {0};
        const double* sp = Sd + S->offsets[b] * (cs ? 2 : 1);
        double* dp = Dd + D->offsets[pos] * (cd ? 2 : 1);
        for (int64_t e = 0; e < tot; ++e) {
            int64_t od = 0;
            for (int i = 0; i < r; ++i) od += idx[i] * dstr[i];
            const double vr = cs ? sp[2 * e] : sp[e], vi = cs ? sp[2 * e + 1] : 0.0;
            const double rr = alpha_re * vr - alpha_im * vi, ri = alpha_re * vi + alpha_im * vr;
            if (cd) { dp[2 * od] += rr; dp[2 * od + 1] += ri; }
            else dp[od] += rr;
            for (int i = 0; i < r; ++i) {
                if (++idx[i] < eS[i]) break;
                idx[i] = 0;
            }
        }
    }
    return 0;
}

/* reference-BLAS dnrm2: scaled sum of squares */
double orc_nrm2(int64_t n, const double* x) {
    double scale = 0.0, ssq = 1.0;
    for (int64_t i = 0; i < n; ++i) {
        if (x[i] != 0.0) {
            const double ax = fabs(x[i]);
            if (scale < ax) { ssq = 1.0 + ssq * (scale / ax) * (scale / ax); scale = ax; }
            else ssq += (ax / scale) * (ax / scale);
        }
    }
    return scale * sqrt(ssq);
}

static int32_t qn_norm(int64_t v, int32_t mod) {
    int64_t m = mod < 0 ? -(int64_t)mod : mod;
    if (m > 1) { int64_t a = v < 0 ? -v : v; return (int32_t)((m * a + v) % m); }
    return (int32_t)v;
}

/* all blocks with sum_j dir_j*qn_j == flux, first index fastest; returns the count */
int64_t orc_flux_blocks(int32_t order, const int32_t* nsect, const int32_t* qn, int32_t nqn, const int32_t* mod,
                        const int32_t* dir, const int32_t* flux, int32_t* blocks, int64_t cap) {
    if (order == 0) return 1;
    int64_t start[ORC_MAXR + 1];
    start[0] = 0;
    for (int j = 0; j < order; ++j) start[j + 1] = start[j] + nsect[j];
    int32_t I[ORC_MAXR] = {0};
    int64_t count = 0;
    for (;;) {
        int ok = 1;
        for (int c = 0; c < nqn && ok; ++c) {
            int32_t acc = 0;
            for (int j = 0; j < order; ++j) {
                int32_t q = qn_norm(qn[(start[j] + I[j]) * nqn + c], mod[c]);
                q = qn_norm((int64_t)q * dir[j], mod[c]);
                acc = qn_norm((int64_t)acc + q, mod[c]);
            }
            if (acc != qn_norm(flux[c], mod[c])) ok = 0;
        }
        if (ok) {
            if (blocks && count < cap) memcpy(blocks + count * order, I, sizeof(int32_t) * (size_t)order);
            ++count;
        }
        int j = 0;
        for (; j < order; ++j) {
            if (++I[j] < nsect[j]) break;
            I[j] = 0;
        }
        if (j == order) break;
    }
    return count;
}
