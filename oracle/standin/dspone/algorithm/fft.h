// TEST INFRASTRUCTURE ONLY (oracle): forwards to the DSPONE/WIPP stand-in.
#ifndef FWD_DSPONE_ALGORITHM_FFT_H
#define FWD_DSPONE_ALGORITHM_FFT_H
#include <dspone/standin.h>
#endif
