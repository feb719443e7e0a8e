/* umbrella header, as the reference's include/mcarray/micarray.h */
#ifndef MCARRAY_B200_MICARRAY_H
#define MCARRAY_B200_MICARRAY_H
#include <mcarray/ArrayDescription.h>
#include <mcarray/ArrayModules.h>
#include <mcarray/BeamformingSeparationAndLocalistaion.h>
#include <mcarray/BinauralLocalisation.h>
#include <mcarray/FastBinauralMasking.h>
#include <mcarray/MultibandBinarualLocalisation.h>
#include <mcarray/SourceLocalisation.h>
#include <mcarray/SourceSeparationAndLocalisation.h>
#endif
