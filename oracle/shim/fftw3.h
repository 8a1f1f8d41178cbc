/* TEST INFRASTRUCTURE -- not part of the product.
 *
 * Minimal FFTW3 API surface so that the *unmodified* Channelflow reference sources
 * (serial build: no HAVE_MPI) compile and link in a container that has no FFTW.
 * Declares exactly the entry points the serial reference calls (flowfield.cpp:577-667,
 * chebyshev.cpp:139-170,262-302,1256-1264, periodicfunc.cpp:96-98).  The implementation is
 * fftw_shim.cpp (our own mixed-radix FFT; results agree with real FFTW to round-off).
 */
#ifndef CF_ORACLE_FFTW3_SHIM_H
#define CF_ORACLE_FFTW3_SHIM_H
#include <stddef.h>
#include <stdio.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef double fftw_complex[2];
struct cf_shim_plan_s;
typedef struct cf_shim_plan_s* fftw_plan;

typedef enum {
    FFTW_R2HC = 0, FFTW_HC2R = 1, FFTW_DHT = 2, FFTW_REDFT00 = 3, FFTW_REDFT01 = 4, FFTW_REDFT10 = 5,
    FFTW_REDFT11 = 6, FFTW_RODFT00 = 7, FFTW_RODFT01 = 8, FFTW_RODFT10 = 9, FFTW_RODFT11 = 10
} fftw_r2r_kind;

#define FFTW_FORWARD (-1)
#define FFTW_BACKWARD (+1)
#define FFTW_MEASURE (0U)
#define FFTW_DESTROY_INPUT (1U << 0)
#define FFTW_UNALIGNED (1U << 1)
#define FFTW_CONSERVE_MEMORY (1U << 2)
#define FFTW_EXHAUSTIVE (1U << 3)
#define FFTW_PRESERVE_INPUT (1U << 4)
#define FFTW_PATIENT (1U << 5)
#define FFTW_ESTIMATE (1U << 6)
#define FFTW_WISDOM_ONLY (1U << 21)

void* fftw_malloc(size_t n);
void fftw_free(void* p);

fftw_plan fftw_plan_many_dft_r2c(int rank, const int* n, int howmany, double* in, const int* inembed, int istride,
                                 int idist, fftw_complex* out, const int* onembed, int ostride, int odist,
                                 unsigned flags);
fftw_plan fftw_plan_many_dft_c2r(int rank, const int* n, int howmany, fftw_complex* in, const int* inembed,
                                 int istride, int idist, double* out, const int* onembed, int ostride, int odist,
                                 unsigned flags);
fftw_plan fftw_plan_r2r_1d(int n, double* in, double* out, fftw_r2r_kind kind, unsigned flags);
fftw_plan fftw_plan_dft_r2c_1d(int n, double* in, fftw_complex* out, unsigned flags);
fftw_plan fftw_plan_dft_c2r_1d(int n, fftw_complex* in, double* out, unsigned flags);

void fftw_execute(const fftw_plan p);
void fftw_execute_r2r(const fftw_plan p, double* in, double* out);
void fftw_destroy_plan(fftw_plan p);

int fftw_import_wisdom_from_file(FILE* f);
void fftw_export_wisdom_to_file(FILE* f);
void fftw_forget_wisdom(void);
void fftw_cleanup(void);

#ifdef __cplusplus
}
#endif
#endif
