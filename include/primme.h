/* primme.h -- umbrella header of the B200-native PRIMME hot-path library.
 *
 * Drop-in for the reference's include/primme.h (reference: include/primme.h:40-123):
 * same integer type, same complex typedef names, same error codes, so that the
 * reference's examples/ and tests/ compile against this tree unchanged.
 * Written from scratch; only the public names/values are shared with the reference.
 */
#ifndef PRIMME_H
#define PRIMME_H

#define PRIMME_VERSION_MAJOR 3
#define PRIMME_VERSION_MINOR 2
/* Marker so callers can tell which implementation they compiled against. */
#define PRIMME_B200_NATIVE 1

/* ---- scalar types that only appear in (unavailable) half/quad entry points ---- */
#if defined(__clang__) && defined(__FLT16_EPSILON__)
#define PRIMME_HALF __fp16
#define PRIMME_WITH_NATIVE_HALF
#else
struct _primme_half {
   int short a;
};
#define PRIMME_HALF struct _primme_half
#endif
#define PRIMME_QUAD double long

struct _primme_complex_half {
   PRIMME_HALF r;
   PRIMME_HALF i;
};
#define PRIMME_COMPLEX_HALF struct _primme_complex_half

#ifdef __cplusplus
#include <complex>
#define PRIMME_COMPLEX_FLOAT std::complex<float>
#define PRIMME_COMPLEX_DOUBLE std::complex<double>
#define PRIMME_COMPLEX_QUAD std::complex<PRIMME_QUAD>
#else
#include <complex.h>
#define PRIMME_COMPLEX_FLOAT float complex
#define PRIMME_COMPLEX_DOUBLE double complex
#define PRIMME_COMPLEX_QUAD long double complex
#endif

/* ---- PRIMME_INT: 64-bit unless the build says otherwise (reference primme.h:82-109) ---- */
#if defined(__cplusplus) && !defined(__STDC_FORMAT_MACROS)
#define __STDC_FORMAT_MACROS
#endif
#include <limits.h>
#include <stdint.h>
#include <inttypes.h>
#if !defined(PRIMME_INT_SIZE) || PRIMME_INT_SIZE == 64
#define PRIMME_INT int64_t
#define PRIMME_INT_P PRId64
#define PRIMME_INT_MAX INT64_MAX
#elif PRIMME_INT_SIZE == 32
#define PRIMME_INT int32_t
#define PRIMME_INT_P PRId32
#define PRIMME_INT_MAX INT32_MAX
#elif PRIMME_INT_SIZE == 0
#define PRIMME_INT int
#define PRIMME_INT_P "d"
#define PRIMME_INT_MAX INT_MAX
#else
#error "unsupported PRIMME_INT_SIZE"
#endif

#include "primme_eigs.h"
#include "primme_svds.h"

/* ---- return codes (reference primme.h:116-123) ---- */
#define PRIMME_UNEXPECTED_FAILURE   (-1)
#define PRIMME_MALLOC_FAILURE       (-2)
#define PRIMME_MAIN_ITER_FAILURE    (-3)
#define PRIMME_LAPACK_FAILURE       (-40)
#define PRIMME_USER_FAILURE         (-41)
#define PRIMME_ORTHO_CONST_FAILURE  (-42)
#define PRIMME_PARALLEL_FAILURE     (-43)
#define PRIMME_FUNCTION_UNAVAILABLE (-44)

#endif /* PRIMME_H */
