/*
 * Plain-C exercise of the C ABI of libodinn_b200.so -- no Python, no ctypes: what a Julia `ccall` (or any FFI) sees.
 *
 *   cc -std=c11 -O1 -I include tests/c_abi/test_capi.c -L odinn.jl_b200/lib -lodinn_b200 -lm -o test_capi
 *   LD_LIBRARY_PATH=odinn.jl_b200/lib ./test_capi
 *
 * Calls odinn_sia2d_rhs / odinn_sia2d_vjp_H / odinn_sia2d_vjp_theta (the replacements of SIA2D_UDE! -> Huginn.SIA2D!,
 * VJP_lambda_dSIAdH and VJP_lambda_dSIAdtheta: src/simulations/inversions/inversion_utils.jl:691-699, src/inverse/SIA2D/VJPs.jl:2-5,
 * 30-33) on a synthetic glacier and checks size-independent properties of the operators:
 *   - dH has a zero border, does not modify H, and conserves mass (sum dH = 0 while no ice touches the border);
 *   - <J v, lambda> == <v, J^T lambda> with J v from central differences of the forward (the reference's own VJP test protocol,
 *     test/SIA2D_adjoint.jl:139-206);
 *   - S == d/dA sum(dH .* lambda) (dH is linear in A when C = 0);
 *   - the error convention: a bad argument returns a negative status and a message, never exits.
 * Exit code 0 on success.  (Values against the oracle are the business of the pytest suite; this file proves the boundary.)
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "odinn_b200.h"

#define PI 3.14159265358979323846
#define NX 70
#define NY 45
#define CHECK(call)                                                                         \
    do {                                                                                    \
        int rc_ = (call);                                                                   \
        if (rc_ != ODINN_OK) {                                                              \
            fprintf(stderr, "%s failed (%d): %s\n", #call, rc_, odinn_last_error(e));       \
            return 1;                                                                       \
        }                                                                                   \
    } while (0)

static double frand(unsigned* s) { /* LCG -> (-1, 1) */
    *s = *s * 1664525u + 1013904223u;
    return ((double)(*s >> 8) / 8388608.0) - 1.0;
}

int main(void) {
    const int nx = NX, ny = NY;
    const double dx = 50.0, A = 2.21e-18; /* test/test_grad_loss.jl:157 */
    static double B[NX * NY], H[NX * NY], Hp[NX * NY], Hm[NX * NY], lam[NX * NY], v[NX * NY];
    static double dH[NX * NY], dHp[NX * NY], dHm[NX * NY], vjp[NX * NY], H_copy[NX * NY];
    unsigned seed = 1234u;
    for (int j = 0; j < ny; ++j)
        for (int i = 0; i < nx; ++i) { /* column-major: element (i, j) at i + j * nx */
            const double x = i * dx, y = j * dx, L = (nx < ny ? nx : ny) * dx;
            const double r = sqrt((x - 0.5 * nx * dx) * (x - 0.5 * nx * dx) + (y - 0.5 * ny * dx) * (y - 0.5 * ny * dx));
            const double cap = 250.0 * (1.0 - (r / (0.35 * L)) * (r / (0.35 * L)));
            B[i + j * nx] = 2000.0 + 0.15 * x + 30.0 * sin(2 * PI * x / 1500.0) * cos(2 * PI * y / 1100.0);
            H[i + j * nx] = cap > 0.0 ? cap : 0.0;
            lam[i + j * nx] = frand(&seed);
            v[i + j * nx] = H[i + j * nx] > 5.0 ? frand(&seed) : 0.0; /* stay away from the H = 0 kink */
        }
    memcpy(H_copy, H, sizeof(H));

    odinn_phys ph = {900.0, 9.81, 1.0, 3.0, 3.0, 0.0, 0.0, 8.5e-20, 8e-17};
    odinn_ensemble* e = NULL;
    int rc = odinn_ensemble_create(0, ODINN_F64, 1, &nx, &ny, &dx, &dx, &ph, &e);
    if (rc != ODINN_OK) {
        fprintf(stderr, "odinn_ensemble_create failed (%d): %s\n", rc, odinn_last_error(NULL));
        return 1;
    }
    CHECK(odinn_upload(e, 0, ODINN_FIELD_B, B, nx));
    CHECK(odinn_set_A_scalar(e, 0, A));

    /* F1 */
    CHECK(odinn_sia2d_rhs(e, 0, H, nx, dH, nx, 0.0));
    if (memcmp(H, H_copy, sizeof(H)) != 0) { fprintf(stderr, "the RHS modified H\n"); return 1; }
    double sum = 0.0, asum = 0.0;
    for (int j = 0; j < ny; ++j)
        for (int i = 0; i < nx; ++i) {
            const double d = dH[i + j * nx];
            if ((i == 0 || j == 0 || i == nx - 1 || j == ny - 1) && d != 0.0) { fprintf(stderr, "nonzero border of dH\n"); return 1; }
            sum += d;
            asum += fabs(d);
        }
    if (!(asum > 0.0) || fabs(sum) > 1e-9 * asum) { fprintf(stderr, "mass not conserved: %g of %g\n", sum, asum); return 1; }

    /* A1: <J v, lambda> == <v, J^T lambda> */
    const double eps = 1e-4;
    for (int k = 0; k < nx * ny; ++k) { Hp[k] = H[k] + eps * v[k]; Hm[k] = H[k] - eps * v[k]; }
    CHECK(odinn_sia2d_rhs(e, 0, Hp, nx, dHp, nx, 0.0));
    CHECK(odinn_sia2d_rhs(e, 0, Hm, nx, dHm, nx, 0.0));
    CHECK(odinn_sia2d_vjp_H(e, 0, lam, nx, H, nx, vjp, nx, 0.0));
    double lhs = 0.0, rhs = 0.0;
    for (int k = 0; k < nx * ny; ++k) {
        lhs += (dHp[k] - dHm[k]) / (2 * eps) * lam[k];
        rhs += vjp[k] * v[k];
        if (!(H[k] > 0.0) && vjp[k] != 0.0) { fprintf(stderr, "VJP nonzero on an ice-free cell\n"); return 1; } /* adjoint.jl:148 */
    }
    if (fabs(lhs - rhs) > 1e-6 * fabs(rhs)) { fprintf(stderr, "transpose identity: %.12g vs %.12g\n", lhs, rhs); return 1; }

    /* A2: S = d/dA sum(dH .* lambda) */
    double S = 0.0, dot0 = 0.0, dot1 = 0.0;
    CHECK(odinn_sia2d_vjp_theta(e, 0, lam, nx, H, nx, &S, 0.0));
    CHECK(odinn_set_A_scalar(e, 0, 2 * A));
    CHECK(odinn_sia2d_rhs(e, 0, H, nx, dHp, nx, 0.0));
    for (int k = 0; k < nx * ny; ++k) { dot0 += dH[k] * lam[k]; dot1 += dHp[k] * lam[k]; }
    if (fabs(S - (dot1 - dot0) / A) > 1e-9 * fabs(S)) { fprintf(stderr, "theta-VJP: %.12g vs %.12g\n", S, (dot1 - dot0) / A); return 1; }

    /* error convention */
    rc = odinn_sia2d_rhs(e, 3, H, nx, dH, nx, 0.0);
    if (rc >= 0 || strlen(odinn_last_error(e)) == 0) { fprintf(stderr, "bad glacier index was accepted\n"); return 1; }
    rc = odinn_sia2d_rhs(e, 0, H, nx - 1, dH, nx, 0.0);
    if (rc >= 0) { fprintf(stderr, "ld < nx was accepted\n"); return 1; }

    printf("test_capi: ok  (launches %lld, <Jv,l> = %.9g, S = %.9g)\n", odinn_launch_count(e), lhs, S);
    odinn_ensemble_destroy(e);
    return 0;
}
