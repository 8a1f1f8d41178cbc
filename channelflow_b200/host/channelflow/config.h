// Build configuration of the B200 host library (stands in for the reference's generated channelflow/config.h):
// serial build, one process per GPU, no FFTW / MPI / NetCDF.
#ifndef CFB200_CONFIG_H
#define CFB200_CONFIG_H
#define CHANNELFLOW_VERSION "2.0-b200"
#define HAVE_DRAND48 1
#define CFB200 1
#endif
