// TEST INFRASTRUCTURE ONLY (oracle): forwards to the DSPONE stand-in.
#ifndef FWD_DSPONE_COMPLEX_H
#define FWD_DSPONE_COMPLEX_H
#include <dspone/standin.h>
#endif
