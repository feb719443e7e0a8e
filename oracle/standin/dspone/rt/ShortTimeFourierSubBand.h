// TEST INFRASTRUCTURE ONLY (oracle): forwards to the DSPONE/WIPP stand-in.
#ifndef FWD_DSPONE_RT_SHORTTIMEFOURIERSUBBAND_H
#define FWD_DSPONE_RT_SHORTTIMEFOURIERSUBBAND_H
#include <dspone/standin.h>
#endif
