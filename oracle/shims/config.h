/* oracle/shims/config.h -- stands in for the meson-generated config.h of the
 * reference (/root/reference/meson.build:44-101) when its hot-path sources are
 * compiled, in place, into oracle/_ref.  TEST INFRASTRUCTURE ONLY. */
#ifndef ORACLE_SHIM_CONFIG_H
#define ORACLE_SHIM_CONFIG_H
#define PACKAGE_TARNAME "polycap"
#define PACKAGE_VERSION "1.2"
#define VERSION "1.2"
#define HAVE_STRDUP 1
#define HAVE_GETTIMEOFDAY 1
#define HAVE_UNLINK 1
#define HAVE_PROPER_COMPLEX_H 1
#define POLYCAP_EXTERN __attribute__((visibility("default"))) extern
#endif
