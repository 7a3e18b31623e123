/* oracle/shims/xraylib.h -- stand-in for the three xraylib calls on the reference's
 * path (src/polycap-photon.c:87-88).  Values come from the same built-in model the
 * product's optical-constants provider uses (polycap_b200/csrc/optconst.c), so that the
 * reference build and the GPU path consume identical amu/scatf.  NOT real xraylib data.
 * TEST INFRASTRUCTURE ONLY. */
#ifndef ORACLE_SHIM_XRAYLIB_H
#define ORACLE_SHIM_XRAYLIB_H
typedef struct xrl_error xrl_error;
double CS_Total(int Z, double E, xrl_error **error);
double Fi(int Z, double E, xrl_error **error);
double AtomicWeight(int Z, xrl_error **error);
#endif
