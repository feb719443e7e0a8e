// TEST INFRASTRUCTURE ONLY (oracle): forwards to the DSPONE/WIPP stand-in.
#ifndef FWD_DSPONE_PF_PREDICTIONMODEL_HPP
#define FWD_DSPONE_PF_PREDICTIONMODEL_HPP
#include <dspone/standin.h>
#endif
