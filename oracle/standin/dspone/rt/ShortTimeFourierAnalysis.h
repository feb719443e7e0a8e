// TEST INFRASTRUCTURE ONLY (oracle): forwards to the DSPONE/WIPP stand-in.
#ifndef FWD_DSPONE_RT_SHORTTIMEFOURIERANALYSIS_H
#define FWD_DSPONE_RT_SHORTTIMEFOURIERANALYSIS_H
#include <dspone/standin.h>
#endif
