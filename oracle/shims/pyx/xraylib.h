/* oracle/shims/pyx/xraylib.h -- stand-in for the xraylib declarations the reference's Cython module pulls in
 * (python/polycap.pyx:43-72: SymbolToAtomicNumber, CompoundParser, error objects).  TEST INFRASTRUCTURE ONLY: lets the
 * UNMODIFIED polycap.pyx be compiled against libpolycap_b200.so in an image without xraylib. */
#ifndef ORACLE_SHIM_PYX_XRAYLIB_H
#define ORACLE_SHIM_PYX_XRAYLIB_H
typedef enum { XRL_ERROR_MEMORY, XRL_ERROR_INVALID_ARGUMENT, XRL_ERROR_IO, XRL_ERROR_TYPE, XRL_ERROR_UNSUPPORTED, XRL_ERROR_RUNTIME } xrl_error_code;
typedef struct { xrl_error_code code; char *message; } xrl_error;
struct compoundData { int nElements; double nAtomsAll; int *Elements; double *massFractions; double *nAtoms; double molarMass; };
int SymbolToAtomicNumber(const char *symbol, xrl_error **error);
void xrlFree(void *);
void xrl_error_free(xrl_error *);
void FreeCompoundData(struct compoundData *cd);
struct compoundData *CompoundParser(const char compoundString[], xrl_error **error);
#endif
