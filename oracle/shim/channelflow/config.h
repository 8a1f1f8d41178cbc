/* TEST INFRASTRUCTURE: stands in for the cmake-generated channelflow/config.h (template config.h.in)
 * when the reference sources are compiled as the CPU oracle.  Serial build: no MPI, NetCDF, HDF5. */
#ifndef CF_ORACLE_CONFIG_H
#define CF_ORACLE_CONFIG_H
#define CHANNELFLOW_VERSION "2.0.2-oracle"
#define COMPILER_VERSION "g++"
#define HAVE_DRAND48
#define HAVE_WORDEXP_H
#endif
