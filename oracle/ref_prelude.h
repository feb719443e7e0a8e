// TEST INFRASTRUCTURE ONLY (oracle).  Force-included (-include) ahead of every translation unit of
// the `make ref` build: pull in the standard headers first, THEN open up the reference classes so
// oracle/ref_capi.cpp can read their per-frame internals without editing the reference sources.
#ifndef ORACLE_REF_PRELUDE_H
#define ORACLE_REF_PRELUDE_H
#include <algorithm>
#include <cmath>
#include <complex>
#include <cstdint>
#include <cstring>
#include <fstream>
#include <functional>
#include <iomanip>
#include <iostream>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <tuple>
#include <utility>
#include <vector>
#include <math.h>
#define private public
#define protected public
#endif
