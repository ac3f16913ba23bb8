// Oracle build shim (test infrastructure only): stands in for the reference's generated config.h.
#ifndef ORACLE_SHIM_CONFIG_H_
#define ORACLE_SHIM_CONFIG_H_
#define HAVE_OMP_H 1
#define PACKAGE "siga"
#define PACKAGE_NAME "siga"
#define PACKAGE_VERSION "oracle"
#define PACKAGE_URL "n/a"
#define PACKAGE_BUGREPORT "n/a"
#endif
