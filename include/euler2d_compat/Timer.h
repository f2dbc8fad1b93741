// src/Timer.h of the reference: `Timer` with start() / stop() / elapsed().
#ifndef EULER2D_COMPAT_TIMER_H
#define EULER2D_COMPAT_TIMER_H
#include "HydroRun.h"
using Timer = euler2d_b200::Timer;
#endif
