/* implementations for oracle/shims/pyx/xraylib.h -- TEST INFRASTRUCTURE ONLY */
#include "xraylib.h"
#include <ctype.h>
#include <stdlib.h>
#include <string.h>
static const char *SYM[] = {"H","He","Li","Be","B","C","N","O","F","Ne","Na","Mg","Al","Si","P","S","Cl","Ar","K","Ca","Sc","Ti","V","Cr","Mn","Fe","Co","Ni","Cu","Zn"};
static const double AW[] = {1.00794,4.002602,6.941,9.012182,10.811,12.0107,14.0067,15.9994,18.9984032,20.1797,22.98977,24.305,26.981538,28.0855,30.973761,32.065,35.453,39.948,39.0983,40.078,44.95591,47.867,50.9415,51.9961,54.938049,55.845,58.9332,58.6934,63.546,65.409};
static xrl_error *mkerr(xrl_error_code c, const char *m) { xrl_error *e = malloc(sizeof *e); e->code = c; e->message = strdup(m); return e; }
int SymbolToAtomicNumber(const char *symbol, xrl_error **error)
{
	for (int i = 0; i < (int)(sizeof SYM / sizeof SYM[0]); i++)
		if (!strcmp(symbol, SYM[i]))
			return i + 1;
	if (error) *error = mkerr(XRL_ERROR_INVALID_ARGUMENT, "Invalid chemical symbol");
	return 0;
}
void xrlFree(void *p) { free(p); }
void xrl_error_free(xrl_error *e) { if (e) { free(e->message); free(e); } }
void FreeCompoundData(struct compoundData *cd) { if (cd) { free(cd->Elements); free(cd->massFractions); free(cd->nAtoms); free(cd); } }
struct compoundData *CompoundParser(const char s[], xrl_error **error)
{ /* flat formulas like SiO2 */
	int z[16]; double n[16]; int k = 0;
	const char *p = s;
	while (*p) {
		char sym[3] = {0, 0, 0};
		if (!isupper((unsigned char)*p)) goto bad;
		sym[0] = *p++;
		if (islower((unsigned char)*p)) sym[1] = *p++;
		int Z = SymbolToAtomicNumber(sym, NULL);
		if (!Z || k == 16) goto bad;
		char *end;
		double c = strtod(p, &end);
		if (end == p) c = 1.0;
		p = end;
		z[k] = Z; n[k] = c; k++;
	}
	if (!k) goto bad;
	struct compoundData *cd = malloc(sizeof *cd);
	cd->nElements = k; cd->Elements = malloc(sizeof(int) * k); cd->massFractions = malloc(sizeof(double) * k); cd->nAtoms = malloc(sizeof(double) * k);
	cd->nAtomsAll = 0; cd->molarMass = 0;
	for (int i = 0; i < k; i++) { cd->Elements[i] = z[i]; cd->nAtoms[i] = n[i]; cd->nAtomsAll += n[i]; cd->molarMass += n[i] * AW[z[i] - 1]; }
	for (int i = 0; i < k; i++) cd->massFractions[i] = n[i] * AW[z[i] - 1] / cd->molarMass;
	return cd;
bad:
	if (error) *error = mkerr(XRL_ERROR_INVALID_ARGUMENT, "composition is not a valid chemical formula");
	return NULL;
}
