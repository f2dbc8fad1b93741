// TEST INFRASTRUCTURE ONLY — function-level known-answer tool: evaluates the UNMODIFIED
// per-cell methods of euler2d::HydroBaseFunctor (/root/reference/src/HydroBaseFunctor.h,
// public and host-callable) on records of doubles read from a file.
//
// usage: ref_kat <file.ini> <func> <in.bin> <out.bin>
//   func    record in (doubles)                                   record out
//   prim    u[ID],u[IP],u[IU],u[IV]                        (4)    q[4], c              (5)
//   slope   q, qPlusX, qMinusX, qPlusY, qMinusY            (20)   dqX[4], dqY[4]       (8)
//   trace   q, dqX, dqY, dtdx, dtdy                        (14)   XMIN,XMAX,YMIN,YMAX  (16)
//   hllc    qleft, qright                                  (8)    flux[4]              (4)
//   approx  qleft, qright                                  (8)    qgdnv[4], flux[4]    (8)
//   cmpflx  qgdnv                                          (4)    flux[4]              (4)
// The .ini supplies gamma0 / smallr / smallc / slope_type exactly as HydroParams::setup
// reads them (through float, config/ConfigMap.cpp:32-40).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "HydroBaseFunctor.h"
#include "HydroParams.h"

using namespace euler2d;

int
main(int argc, char * argv[])
{
  if (argc != 5)
  {
    fprintf(stderr, "usage: ref_kat <file.ini> <func> <in.bin> <out.bin>\n");
    return 2;
  }
  ConfigMap   configMap(argv[1]);
  HydroParams params = HydroParams();
  params.setup(configMap);
  HydroBaseFunctor  fn(params);
  const std::string func = argv[2];

  int nin = 0, nout = 0;
  if (func == "prim")
    nin = 4, nout = 5;
  else if (func == "slope")
    nin = 20, nout = 8;
  else if (func == "trace")
    nin = 14, nout = 16;
  else if (func == "hllc")
    nin = 8, nout = 4;
  else if (func == "approx")
    nin = 8, nout = 8;
  else if (func == "cmpflx")
    nin = 4, nout = 4;
  else
  {
    fprintf(stderr, "ref_kat: unknown func %s\n", func.c_str());
    return 2;
  }

  FILE * fi = fopen(argv[3], "rb");
  if (!fi)
    return 2;
  fseek(fi, 0, SEEK_END);
  long bytes = ftell(fi);
  fseek(fi, 0, SEEK_SET);
  size_t              nrec = bytes / (sizeof(double) * nin);
  std::vector<double> in(nrec * nin), out(nrec * nout);
  if (fread(in.data(), sizeof(double), in.size(), fi) != in.size())
    return 2;
  fclose(fi);

  auto load = [](const double * p) {
    HydroState s;
    for (int k = 0; k < 4; ++k)
      s[k] = p[k];
    return s;
  };
  auto store = [](double * p, const HydroState & s) {
    for (int k = 0; k < 4; ++k)
      p[k] = s[k];
  };

  for (size_t r = 0; r < nrec; ++r)
  {
    const double * a = &in[r * nin];
    double *       o = &out[r * nout];
    if (func == "prim")
    {
      HydroState q;
      real_t     c = 0;
      fn.computePrimitives(load(a), &c, q);
      store(o, q);
      o[4] = c;
    }
    else if (func == "slope")
    {
      HydroState dqX, dqY;
      for (int k = 0; k < 4; ++k)
        dqX[k] = dqY[k] = 0;
      fn.slope_unsplit_hydro_2d(load(a), load(a + 4), load(a + 8), load(a + 12), load(a + 16), dqX, dqY);
      store(o, dqX);
      store(o + 4, dqY);
    }
    else if (func == "trace")
    {
      HydroState q = load(a), dqX = load(a + 4), dqY = load(a + 8), qf;
      for (int face = 0; face < 4; ++face)
      {
        fn.trace_unsplit_2d_along_dir(q, dqX, dqY, a[12], a[13], face, qf);
        store(o + 4 * face, qf);
      }
    }
    else if (func == "hllc")
    {
      HydroState qg, fl;
      fn.riemann_hllc(load(a), load(a + 4), qg, fl);
      store(o, fl);
    }
    else if (func == "approx")
    {
      HydroState qg, fl;
      fn.riemann_approx(load(a), load(a + 4), qg, fl);
      store(o, qg);
      store(o + 4, fl);
    }
    else if (func == "cmpflx")
    {
      HydroState fl;
      fn.cmpflx(load(a), fl);
      store(o, fl);
    }
  }
  FILE * fo = fopen(argv[4], "wb");
  if (!fo)
    return 2;
  fwrite(out.data(), sizeof(double), out.size(), fo);
  fclose(fo);
  return 0;
}
