// TEST INFRASTRUCTURE ONLY (oracle): forwards to the DSPONE/WIPP stand-in.
#ifndef FWD_WIPP_WIPPUTILS_H
#define FWD_WIPP_WIPPUTILS_H
#include <wipp/wipp.h>
#endif
