/* Minimal stand-in for PyTorch-0.2's TH.h, TEST INFRASTRUCTURE ONLY.
 *
 * The reference CPU implementation (/root/reference/my_package/src/my_lib.c)
 * touches TH through exactly three things: the `size[]` / `stride[]` members
 * of THFloatTensor and THFloatTensor_data().  This header provides just that
 * so the reference file compiles UNCHANGED, from where it lies, into
 * oracle/_ref/libmemc_ref_cpu.so (see oracle/Makefile).  my_lib.c includes
 * <TH.h> twice (my_lib.c:1 and :893), hence the guard.
 */
#ifndef MEMC_ORACLE_TH_STUB_H
#define MEMC_ORACLE_TH_STUB_H

typedef struct THFloatTensor {
    long *size;
    long *stride;
    int nDimension;
    float *data;
} THFloatTensor;

static inline float *THFloatTensor_data(const THFloatTensor *t) { return t->data; }

#endif
