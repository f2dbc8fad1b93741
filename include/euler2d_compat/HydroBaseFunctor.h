// src/HydroBaseFunctor.h of the reference holds device code only (computePrimitives, slopes, trace, Riemann solvers);
// its B200 counterparts are the kernels inside libeuler2d_b200.so, so a host program needs nothing from here.
