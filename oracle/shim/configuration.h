/* configuration.h -- stands in for the file CMake generates from fastcard/configuration.h.in */
#ifndef THR_ORACLE_SHIM_CONFIGURATION_H
#define THR_ORACLE_SHIM_CONFIGURATION_H
#define VERSION_MAJOR 0
#define VERSION_MINOR 10
#define VERSION_STRING "0.10"
#ifndef USE_FFTW
#define USE_FFTW
#endif
#endif
