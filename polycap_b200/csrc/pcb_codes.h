/* pcb_codes.h -- row numbers of the per-energy constant table and the status codes of a finished photon (device and host). */
#ifndef PCB_CODES_H
#define PCB_CODES_H

/* energy-constant rows (one double per energy each) */
enum { EC_NRE = 0, EC_NIM, EC_I2RE, EC_I2IM, EC_KR, EC_AMU, EC_ROWS };

/* status codes of a finished photon: the reference's launch return codes (photon.c:947-954)
 * plus 3 = "returned 1 but missed the exit window" (source.c:762-777). */
enum { ST_EXIT = 1, ST_ABSORBED = 0, ST_NOT_ENTERED = 2, ST_OUTSIDE = -2, ST_ERROR = -1, ST_MISSED_WINDOW = 3 };

#endif
