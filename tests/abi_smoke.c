/* Plain-C driver of include/qrochet_b200.h: proves the header with a C compiler (not with ctypes prototypes) and walks
 * the boundary a Julia `ccall` takes: create -> tensors -> contract -> MPS -> canonize! -> evolve! -> overlap -> expect.
 * Built by build.sh (gcc -std=c99 -Wall -Werror), run on a GPU by tests/test_gpu_abi.py.
 *   usage: abi_smoke            run on device 0, print "ABI_SMOKE_OK" and exit 0
 *          abi_smoke --symbols  only check that the library resolves (no GPU needed) */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "qrochet_b200.h"

#define CHECK(call)                                                                               \
    do {                                                                                          \
        if (verbose) {                                                                            \
            fprintf(stderr, "[abi_smoke] %s\n", #call);                                           \
            fflush(stderr);                                                                       \
        }                                                                                         \
        int32_t rc_ = (call);                                                                     \
        if (rc_ != QB200_OK) {                                                                    \
            fprintf(stderr, "%s:%d: %s -> %d (%s)\n", __FILE__, __LINE__, #call, (int)rc_,        \
                    qb200_last_error(ctx));                                                       \
            return 1;                                                                             \
        }                                                                                         \
    } while (0)

static int verbose = 0;

static double frand(unsigned* s) {
    *s = *s * 1664525u + 1013904223u;
    return ((*s >> 8) & 0xffffff) / (double)0x1000000 - 0.5;
}

int main(int argc, char** argv) {
    qb200_ctx* ctx = NULL;
    if (argc > 1 && strcmp(argv[1], "--symbols") == 0) {
        /* taking the addresses forces the dynamic linker to resolve every symbol used below */
        void* fns[] = {(void*)qb200_create, (void*)qb200_contract, (void*)qb200_mps_evolve2, (void*)qb200_mps_overlap,
                       (void*)qb200_mps_expect, (void*)qb200_svd, (void*)qb200_qr, (void*)qb200_tn_plan};
        printf("ABI_SYMBOLS_OK %d\n", (int)(sizeof(fns) / sizeof(fns[0])));
        return 0;
    }
    verbose = getenv("ABI_SMOKE_VERBOSE") != NULL;
    if (qb200_create(0, &ctx) != QB200_OK) {
        fprintf(stderr, "qb200_create failed: %s\n", qb200_last_error(NULL));
        return 2;
    }
    unsigned seed = 12345u;

    /* ---- Tenet level: C[i,k] = sum_j A[i,j] B[j,k] with int32 mode labels, checked on the host ---- */
    enum { M = 5, K = 7, N = 3 };
    double a[2 * M * K], b[2 * K * N], c[2 * M * N];
    for (int i = 0; i < 2 * M * K; ++i) a[i] = frand(&seed);
    for (int i = 0; i < 2 * K * N; ++i) b[i] = frand(&seed);
    int64_t ea[2] = {M, K}, eb[2] = {K, N}, ec[2] = {M, N};
    qb200_tensor *ta, *tb, *tc;
    CHECK(qb200_tensor_alloc(ctx, QB200_C128, 2, ea, &ta));
    CHECK(qb200_tensor_alloc(ctx, QB200_C128, 2, eb, &tb));
    CHECK(qb200_tensor_alloc(ctx, QB200_C128, 2, ec, &tc));
    CHECK(qb200_tensor_upload(ctx, ta, a));
    CHECK(qb200_tensor_upload(ctx, tb, b));
    int32_t ma[2] = {10, 11}, mb[2] = {11, 12}, mc[2] = {10, 12};
    CHECK(qb200_contract(ctx, ta, ma, 0, tb, mb, 0, tc, mc, NULL, NULL));
    CHECK(qb200_tensor_download(ctx, tc, c));
    double worst = 0.0;
    for (int i = 0; i < M; ++i)
        for (int k = 0; k < N; ++k) {
            double re = 0.0, im = 0.0;
            for (int j = 0; j < K; ++j) {
                double ar = a[2 * (i + M * j)], ai = a[2 * (i + M * j) + 1];
                double br = b[2 * (j + K * k)], bi = b[2 * (j + K * k) + 1];
                re += ar * br - ai * bi;
                im += ar * bi + ai * br;
            }
            worst = fmax(worst, fabs(re - c[2 * (i + M * k)]) + fabs(im - c[2 * (i + M * k) + 1]));
        }
    if (worst > 1e-13) {
        fprintf(stderr, "contract: max error %.3e\n", worst);
        return 1;
    }
    CHECK(qb200_tensor_free(ctx, ta));
    CHECK(qb200_tensor_free(ctx, tb));
    CHECK(qb200_tensor_free(ctx, tc));

    /* ---- chain level: 6-site product-like random MPS (bond 4) -> canonize! -> evolve! -> overlap / expect ---- */
    enum { NS = 6 };
    int64_t chi[NS + 1] = {1, 2, 4, 4, 4, 2, 1};
    qb200_mps *psi, *phi;
    CHECK(qb200_mps_create(ctx, NS, &psi));
    for (int s = 0; s < NS; ++s) {
        int64_t cnt = chi[s] * 2 * chi[s + 1];
        double* buf = (double*)malloc(sizeof(double) * 2 * (size_t)cnt);
        for (int64_t i = 0; i < 2 * cnt; ++i) buf[i] = frand(&seed);
        CHECK(qb200_mps_set_site(ctx, psi, s, chi[s], 2, chi[s + 1], buf));
        free(buf);
    }
    double n0[2], n1[2], ov[2], ex[2];
    CHECK(qb200_mps_overlap(ctx, psi, psi, n0));
    CHECK(qb200_mps_copy(ctx, psi, &phi));
    CHECK(qb200_mps_canonize(ctx, psi));
    if (qb200_mps_form(psi) != 1) {
        fprintf(stderr, "canonize!: form %d\n", (int)qb200_mps_form(psi));
        return 1;
    }
    CHECK(qb200_mps_overlap(ctx, psi, psi, n1));
    CHECK(qb200_mps_overlap(ctx, psi, phi, ov));
    if (fabs(n1[0] - n0[0]) > 1e-10 * n0[0] || fabs(ov[0] - n0[0]) > 1e-10 * n0[0] || fabs(ov[1]) > 1e-10 * n0[0]) {
        fprintf(stderr, "canonize! changed the state: %.15g %.15g (%.15g, %.15g)\n", n0[0], n1[0], ov[0], ov[1]);
        return 1;
    }
    /* sum lambda^2 = |psi|^2 on the middle bond (test/Ansatz/Chain_test.jl:322-323) */
    double lam[8];
    int64_t nl = 0;
    CHECK(qb200_mps_get_lambda(ctx, psi, 2, lam, &nl));
    double s2 = 0.0;
    for (int64_t i = 0; i < nl; ++i) s2 += lam[i] * lam[i];
    if (nl != 4 || fabs(s2 - n0[0]) > 1e-10 * n0[0]) {
        fprintf(stderr, "Schmidt vector: %d values, sum^2 %.15g vs %.15g\n", (int)nl, s2, n0[0]);
        return 1;
    }
    /* identity gate, Vidal branch, threshold 1e-10 (drops the numerically zero half of theta's spectrum): Schmidt values unchanged; CNOT-like permutation gate: norm kept */
    double gate[32];
    memset(gate, 0, sizeof(gate));
    for (int i = 0; i < 4; ++i) gate[2 * (i + 4 * i)] = 1.0;
    int64_t kept = 0;
    double dw = 0.0, lam2[8];
    CHECK(qb200_mps_evolve2(ctx, psi, 2, gate, 0, 1e-10, 0, 1, &kept, &dw));
    CHECK(qb200_mps_get_lambda(ctx, psi, 2, lam2, &nl));
    for (int64_t i = 0; i < nl; ++i)
        if (fabs(lam2[i] - lam[i]) > 1e-12 * lam[0]) {
            fprintf(stderr, "identity gate moved Schmidt value %d: %.15g -> %.15g\n", (int)i, lam[i], lam2[i]);
            return 1;
        }
    if (kept != 4 || dw > 1e-20) {
        fprintf(stderr, "identity gate: kept %d, discarded %.3e\n", (int)kept, dw);
        return 1;
    }
    memset(gate, 0, sizeof(gate));
    { /* swap |01> <-> |10> : a unitary permutation */
        int perm[4] = {0, 2, 1, 3};
        for (int i = 0; i < 4; ++i) gate[2 * (perm[i] + 4 * i)] = 1.0;
    }
    CHECK(qb200_mps_evolve2(ctx, psi, 3, gate, 4, -1.0, 1, 1, &kept, &dw));
    /* expect(psi, [I_3]) = |psi|^2 through the general-observable entry point (Chain.jl:724-735) */
    double eye[8] = {1, 0, 0, 0, 0, 0, 1, 0};
    int32_t nlanes[1] = {1}, sites[1] = {3};
    CHECK(qb200_mps_expect(ctx, psi, 1, nlanes, sites, eye, ex));
    CHECK(qb200_mps_overlap(ctx, psi, psi, n1));
    if (fabs(ex[0] - n1[0]) > 1e-10 * n1[0] || fabs(ex[1]) > 1e-10 * n1[0]) {
        fprintf(stderr, "expect(psi, [I]) = (%.15g, %.15g) vs |psi|^2 = %.15g\n", ex[0], ex[1], n1[0]);
        return 1;
    }
    /* error path: bond out of range returns QB200_E_INVALID and a message, never aborts */
    if (qb200_mps_evolve2(ctx, psi, NS, gate, 0, -1.0, 0, 1, NULL, NULL) != QB200_E_INVALID || !*qb200_last_error(ctx)) {
        fprintf(stderr, "missing QB200_E_INVALID for an out-of-range bond\n");
        return 1;
    }
    CHECK(qb200_mps_free(ctx, psi));
    CHECK(qb200_mps_free(ctx, phi));
    CHECK(qb200_synchronize(ctx));
    printf("ABI_SMOKE_OK launches=%lld\n", (long long)qb200_launch_count(ctx));
    qb200_destroy(ctx);
    return 0;
}
