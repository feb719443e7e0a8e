// TEST INFRASTRUCTURE ONLY (oracle): forwards to the DSPONE/WIPP stand-in.
#ifndef FWD_DSPONE_FILTER_FILTERBANKMELSCALE_H
#define FWD_DSPONE_FILTER_FILTERBANKMELSCALE_H
#include <dspone/standin.h>
#endif
