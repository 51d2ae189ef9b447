// tfhe_min.h -- the handful of TFHE data types the imputation path touches, as plain data holders, for
// building the host layer where the TFHE library is not installed (the GPU box). Field names, order and
// types follow tfhe/src/include/polynomials.h:10-34, tlwe.h:9-60 so that code written against <tfhe.h>
// reads the same members (the structs themselves are plain: no constructors, no ownership).
// No TFHE arithmetic lives here: on this path all of it runs in libidash_b200.so.
#ifndef IDASH_B200_TFHE_MIN_H
#define IDASH_B200_TFHE_MIN_H

#include <cstdint>

typedef int32_t Torus32;   // tfhe_core.h: the torus R/Z scaled by 2^32

struct IntPolynomial { int32_t N; int32_t *coefs; };
struct TorusPolynomial { int32_t N; Torus32 *coefsT; };

struct TLweParams {
    int32_t N, k;
    double alpha_min, alpha_max;
};

struct TLweKey {
    const TLweParams *params;
    IntPolynomial *key;          // k binary polynomials
};

struct TLweSample {
    TorusPolynomial *a;          // k + 1 polynomials: mask, then right-hand side
    TorusPolynomial *b;          // = a + k
    double current_variance;
    int32_t k;
};

#endif
