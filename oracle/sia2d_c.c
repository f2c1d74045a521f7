/*
 * oracle/sia2d_c.c -- plain-C restatement of ODINN.jl's SIA2D forward + discrete VJPs.
 * TEST INFRASTRUCTURE / TIMED CPU BASELINE ONLY (never linked into the product library).
 * PARITY UNPINNED BY STORED NUMBERS: see oracle/__init__.py.  The arithmetic is in
 * sia2d_c_impl.h, instantiated here for Float64 (the reference default) and Float32
 * (Sleipnir.doublePrec = false build, test/SIA2D_adjoint_utils.jl:22).
 * Build: gcc -O3 -march=x86-64-v3 -fopenmp -fPIC -shared sia2d_c.c -lm  (see __graft_entry__.build_oracle).
 */
#include <math.h>
#include <stddef.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define REAL double
#define SUF _f64
#include "sia2d_c_impl.h"
#undef REAL
#undef SUF

#define REAL float
#define SUF _f32
#include "sia2d_c_impl.h"
#undef REAL
#undef SUF

/* Use n OpenMP threads from now on (n <= 0: leave unchanged).  The timed CPU arm calls this with the number of host
 * cores available to the process: launchers such as torchrun export OMP_NUM_THREADS=1, which would otherwise cripple
 * the baseline and inflate every GPU/CPU ratio. */
void sia2d_oracle_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int sia2d_oracle_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
