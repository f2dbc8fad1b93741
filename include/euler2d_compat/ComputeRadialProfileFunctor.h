// src/ComputeRadialProfileFunctor.h of the reference, served by the B200 library (everything lives in HydroRun.hpp).
#include "HydroRun.h"
