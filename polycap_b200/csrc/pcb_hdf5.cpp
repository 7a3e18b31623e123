/* pcb_hdf5.cpp -- HDF5 writer (placeholder): reports that the writer is not built yet. */
#include <cstdio>
#include "pcb_internal.h"
extern "C" int pcb_hdf5_write(polycap_transmission_efficiencies *, const char *, char *errbuf, size_t errlen)
{
	snprintf(errbuf, errlen, "HDF5 output is not available in this build");
	return -1;
}
