// .ini -> e2d_params, with the reference's exact reading semantics.
//
// Replaces config/inih/ini.cpp + INIReader.cpp (the parser), config/ConfigMap.cpp (typed getters) and
// HydroParams::setup / init / print (src/HydroParams.cpp:43-224) of the reference.  What has to be
// preserved for bit parity of everything downstream:
//   * every real-valued key is parsed with strtof and widened to double (ConfigMap.cpp:32-40), and the
//     defaults are float arguments too: gamma0=1.666 -> 0x1.aa7efap+0, smallr default 1e-10f, ...
//   * integers go through strtol(base 0) (INIReader.cpp:57-66); implementationVersion through strtof
//   * keys are "section.name" lower-cased (INIReader.cpp:93-100), last assignment wins
//   * ';' starts an inline comment only after whitespace, '#' only at line start, lines are cut at
//     199 characters, indented lines continue (replace) the previous value (ini.cpp:42-52,87-144)
//   * a missing file is NOT an error for the reference (ParseError() is never checked, main.cpp:76);
//     here it returns E2D_ERR_IO after filling the defaults so the caller can decide.
#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <new>
#include <sstream>
#include <string>

#include "../../include/euler2d_b200.h"

namespace
{

using KeyMap = std::map<std::string, std::string>;

std::string
make_key(const std::string & section, const std::string & name)
{
  std::string key = section + "." + name;
  for (auto & ch : key)
    ch = static_cast<char>(std::tolower(static_cast<unsigned char>(ch)));
  return key;
}

char *
strip_right(char * s)
{
  char * p = s + std::strlen(s);
  while (p > s && std::isspace(static_cast<unsigned char>(*--p)))
    *p = '\0';
  return s;
}

char *
skip_left(char * s)
{
  while (*s && std::isspace(static_cast<unsigned char>(*s)))
    ++s;
  return s;
}

// first occurrence of `c`, or of a ';' preceded by whitespace, or the terminating NUL
char *
scan_to(char * s, char c)
{
  bool prev_space = false;
  for (; *s && *s != c && !(prev_space && *s == ';'); ++s)
    prev_space = std::isspace(static_cast<unsigned char>(*s)) != 0;
  return s;
}

constexpr int kMaxLine = 200, kMaxSection = 50, kMaxName = 50;

// Line-oriented parse; `next_line` yields successive chunks the way fgets(line, 200, f) would.
template <class NextLine>
void
parse_ini(NextLine next_line, KeyMap & kv)
{
  char        line[kMaxLine];
  std::string section, prev_name;
  while (next_line(line, kMaxLine))
  {
    char * start = skip_left(strip_right(line));
    if (!prev_name.empty() && *start && start > line)
    {
      kv[make_key(section, prev_name)] = start; // continuation line
      continue;
    }
    if (*start == ';' || *start == '#' || *start == '\0')
      continue;
    if (*start == '[')
    {
      char * end = scan_to(start + 1, ']');
      if (*end == ']')
      {
        *end = '\0';
        section.assign(start + 1, strnlen(start + 1, kMaxSection - 1));
        prev_name.clear();
      }
      continue;
    }
    char * end = scan_to(start, '=');
    if (*end != '=')
      continue; // the reference records a parse error and ignores the line
    *end = '\0';
    char * name = strip_right(start);
    char * value = skip_left(end + 1);
    end = scan_to(value, '\0');
    if (*end == ';')
      *end = '\0';
    strip_right(value);
    prev_name.assign(name, strnlen(name, kMaxName - 1));
    kv[make_key(section, name)] = value;
  }
}

struct Config
{
  KeyMap kv;

  const std::string *
  find(const char * section, const char * name) const
  {
    auto it = kv.find(make_key(section, name));
    return it == kv.end() ? nullptr : &it->second;
  }
  long
  integer(const char * section, const char * name, long dflt) const
  {
    const std::string * v = find(section, name);
    const char *        str = v ? v->c_str() : "";
    char *              end;
    long                n = std::strtol(str, &end, 0);
    return end > str ? n : dflt;
  }
  float
  real(const char * section, const char * name, float dflt) const
  {
    const std::string * v = find(section, name);
    const char *        str = v ? v->c_str() : "";
    char *              end;
    float               x = std::strtof(str, &end);
    return end > str ? x : dflt;
  }
  bool
  boolean(const char * section, const char * name, bool dflt) const
  {
    const std::string * v = find(section, name);
    if (!v || v->empty())
      return dflt;
    if (*v == "1" || *v == "yes" || *v == "true" || *v == "on")
      return true;
    if (*v == "0" || *v == "no" || *v == "false" || *v == "off")
      return false;
    return dflt;
  }
  std::string
  string(const char * section, const char * name, const char * dflt) const
  {
    const std::string * v = find(section, name);
    return v ? *v : std::string(dflt);
  }
};

void
copy_string(char (&dst)[256], const std::string & src)
{
  std::snprintf(dst, sizeof dst, "%s", src.c_str());
}

void
setup(const Config & cfg, e2d_params * p)
{
  std::memset(p, 0, sizeof *p);
  p->enableOutput = 1;
  p->ghostWidth = 2;

  // [run]
  p->nStepmax = static_cast<int>(cfg.integer("run", "nstepmax", 1000));
  p->tEnd = cfg.real("run", "tend", 0.0);
  p->nOutput = static_cast<int>(cfg.integer("run", "noutput", 100));
  if (p->nOutput == -1)
    p->enableOutput = 0;
  p->ioHDF5 = cfg.boolean("run", "use_HDF5", false); // built without HDF5, like the oracle
  p->ioVTK = cfg.boolean("run", "use_VTK", true);

  // [mesh]
  p->nx = static_cast<int>(cfg.integer("mesh", "nx", 2));
  p->ny = static_cast<int>(cfg.integer("mesh", "ny", 2));
  p->xmin = cfg.real("mesh", "xmin", 0.0);
  p->ymin = cfg.real("mesh", "ymin", 0.0);
  p->xmax = cfg.real("mesh", "xmax", 1.0);
  p->ymax = cfg.real("mesh", "ymax", 1.0);
  p->boundary_type_xmin = static_cast<int>(cfg.integer("mesh", "boundary_type_xmin", E2D_BC_DIRICHLET));
  p->boundary_type_xmax = static_cast<int>(cfg.integer("mesh", "boundary_type_xmax", E2D_BC_DIRICHLET));
  p->boundary_type_ymin = static_cast<int>(cfg.integer("mesh", "boundary_type_ymin", E2D_BC_DIRICHLET));
  p->boundary_type_ymax = static_cast<int>(cfg.integer("mesh", "boundary_type_ymax", E2D_BC_DIRICHLET));

  // [hydro]
  p->gamma0 = cfg.real("hydro", "gamma0", 1.4);
  p->cfl = cfg.real("hydro", "cfl", 0.5);
  p->slope_type = cfg.real("hydro", "slope_type", 1.0);
  p->smallc = cfg.real("hydro", "smallc", 1e-10);
  p->smallr = cfg.real("hydro", "smallr", 1e-10);
  p->niter_riemann = static_cast<int>(cfg.integer("hydro", "niter_riemann", 10));

  const std::string riemann = cfg.string("hydro", "riemann", "approx");
  if (riemann == "approx")
    p->riemannSolverType = E2D_RIEMANN_APPROX;
  else if (riemann == "hll")
    p->riemannSolverType = E2D_RIEMANN_HLL;
  else if (riemann == "hllc")
    p->riemannSolverType = E2D_RIEMANN_HLLC;
  else if (riemann == "rusanov" || riemann == "llf") // extension (the reference would print the message below)
    p->riemannSolverType = E2D_RIEMANN_RUSANOV;
  else
  {
    std::printf("Riemann Solver specified in parameter file is invalid\n");
    std::printf("Use the default one : approx\n");
    p->riemannSolverType = E2D_RIEMANN_APPROX;
  }

  const std::string problem = cfg.string("hydro", "problem", "unknown");
  if (problem == "implode")
    p->problemType = E2D_PROBLEM_IMPLODE;
  else if (problem == "blast")
    p->problemType = E2D_PROBLEM_BLAST;
  else if (problem == "four_quadrant")
    p->problemType = E2D_PROBLEM_FOUR_QUADRANT;
  else if (problem == "discontinuity")
    p->problemType = E2D_PROBLEM_DISCONTINUITY;
  else if (problem == "shocked_bubble")
  {
    p->problemType = E2D_PROBLEM_SHOCKED_BUBBLE;
    p->bubble_radius = cfg.real("shocked_bubble", "bubble_radius", 0.025);
    p->bubble_center_x = cfg.real("shocked_bubble", "bubble_center_x", 0.225);
    p->bubble_center_y = cfg.real("shocked_bubble", "bubble_center_y", 0.0445);
    p->bubble_density = cfg.real("shocked_bubble", "bubble_density", 3.863);
    p->bubble_pressure = cfg.real("shocked_bubble", "bubble_pressure", 1.0132e5);
    p->preshock_density = cfg.real("shocked_bubble", "preshock_density", 1.225);
    p->preshock_pressure = cfg.real("shocked_bubble", "preshock_pressure", 1.0132e5);
    p->postshock_density = cfg.real("shocked_bubble", "postshock_density", 1.686);
    p->postshock_pressure = cfg.real("shocked_bubble", "postshock_pressure", 1.59e5);
    p->postshock_velocity = cfg.real("shocked_bubble", "postshock_velocity", 113.5);
    p->shock_loc = cfg.real("shocked_bubble", "shock_loc", 0.170);
  }
  else
  {
    std::printf("Problem is invalid\n");
    std::printf("Use the default one : implode\n");
    p->problemType = E2D_PROBLEM_IMPLODE;
  }

  // [blast] — the fallback values are formed in double and then narrowed by the float parameter
  p->blast_radius = cfg.real("blast", "radius", static_cast<float>((p->xmin + p->xmax) / 2.0 / 10));
  p->blast_center_x = cfg.real("blast", "center_x", static_cast<float>((p->xmin + p->xmax) / 2));
  p->blast_center_y = cfg.real("blast", "center_y", static_cast<float>((p->ymin + p->ymax) / 2));
  p->blast_density_in = cfg.real("blast", "density_in", 1.0);
  p->blast_density_out = cfg.real("blast", "density_out", 1.2);
  p->blast_pressure_in = cfg.real("blast", "pressure_in", 10.0);
  p->blast_pressure_out = cfg.real("blast", "pressure_out", 0.1);
  p->blast_total_energy_inside = cfg.real("blast", "total_energy_inside", 0.0);
  p->blast_nbins = static_cast<int>(cfg.integer("blast", "nbins", 100));

  // [other]
  p->implementationVersion = static_cast<int>(cfg.real("OTHER", "implementationVersion", 0));
  if (p->implementationVersion != 0 && p->implementationVersion != 1 && p->implementationVersion != 2)
  {
    std::printf("Implementation version is invalid (must be 0, 1 or 2)\n");
    std::printf("Use the default : 0\n");
    p->implementationVersion = 0;
  }
  p->honourRiemannSolver = cfg.boolean("OTHER", "honourRiemannSolver", false) ? 1 : 0;
  p->unfusedKernels = cfg.boolean("OTHER", "unfusedKernels", false) ? 1 : 0;
  {
    const std::string ar = cfg.string("OTHER", "arithmetic", "strict");
    p->arithmetic = (ar == "fast") ? E2D_ARITH_FAST : E2D_ARITH_STRICT;
  }

  // [output]
  p->vtkAppended = cfg.boolean("output", "vtk_appended", false) ? 1 : 0;
  copy_string(p->outputDir, cfg.string("output", "outputDir", "./"));
  copy_string(p->outputPrefix, cfg.string("output", "outputPrefix", "output"));

  e2d_params_init(p);
}

} // namespace

extern "C" int
e2d_params_init(e2d_params * p)
{
  if (!p)
    return E2D_ERR_INVALID;
  p->imin = 0;
  p->jmin = 0;
  p->imax = p->nx - 1 + 2 * p->ghostWidth;
  p->jmax = p->ny - 1 + 2 * p->ghostWidth;
  p->isize = p->imax - p->imin + 1;
  p->jsize = p->jmax - p->jmin + 1;
  p->dx = (p->xmax - p->xmin) / p->nx;
  p->dy = (p->ymax - p->ymin) / p->ny;
  p->smallp = p->smallc * p->smallc / p->gamma0;
  p->smallpp = p->smallr * p->smallp;
  p->gamma6 = (p->gamma0 + 1.0) / (2.0 * p->gamma0);
  return E2D_OK;
}

// ---- ConfigMap over the C ABI (config/ConfigMap.h:26-46, config/inih/INIReader.h) ----
struct e2d_config
{
  Config cfg;
  int    parse_error = 0; // INIReader::ParseError(): -1 when the file could not be opened
};

extern "C" int
e2d_config_open(const char * path, e2d_config ** out)
{
  if (!path || !out)
    return E2D_ERR_INVALID;
  e2d_config * c = new (std::nothrow) e2d_config();
  if (!c)
    return E2D_ERR_ALLOC;
  FILE * f = std::fopen(path, "r");
  if (f)
  {
    parse_ini([f](char * buf, int n) { return std::fgets(buf, n, f) != nullptr; }, c->cfg.kv);
    std::fclose(f);
  }
  else
    c->parse_error = -1;
  *out = c;
  return f ? E2D_OK : E2D_ERR_IO; // the handle is valid either way (an empty map), like the reference's object
}

extern "C" int
e2d_config_from_string(const char * ini_text, e2d_config ** out)
{
  if (!ini_text || !out)
    return E2D_ERR_INVALID;
  e2d_config * c = new (std::nothrow) e2d_config();
  if (!c)
    return E2D_ERR_ALLOC;
  const char * cur = ini_text;
  parse_ini(
    [&cur](char * buf, int n) {
      if (!*cur)
        return false;
      int k = 0;
      while (k < n - 1 && *cur)
      {
        buf[k++] = *cur;
        if (*cur++ == '\n')
          break;
      }
      buf[k] = '\0';
      return true;
    },
    c->cfg.kv);
  *out = c;
  return E2D_OK;
}

extern "C" void
e2d_config_close(e2d_config * c)
{
  delete c;
}

extern "C" int
e2d_config_parse_error(const e2d_config * c)
{
  return c ? c->parse_error : -1;
}

extern "C" float
e2d_config_get_float(const e2d_config * c, const char * section, const char * name, float dflt)
{
  return (c && section && name) ? c->cfg.real(section, name, dflt) : dflt;
}

extern "C" long
e2d_config_get_integer(const e2d_config * c, const char * section, const char * name, long dflt)
{
  return (c && section && name) ? c->cfg.integer(section, name, dflt) : dflt;
}

extern "C" int
e2d_config_get_bool(const e2d_config * c, const char * section, const char * name, int dflt)
{
  return (c && section && name) ? (c->cfg.boolean(section, name, dflt != 0) ? 1 : 0) : dflt;
}

extern "C" int
e2d_config_get_string(const e2d_config * c, const char * section, const char * name, const char * dflt, char * buf,
                      size_t cap)
{
  if (!buf || cap == 0)
    return E2D_ERR_INVALID;
  const std::string v = (c && section && name) ? c->cfg.string(section, name, dflt ? dflt : "") : std::string(dflt ? dflt : "");
  std::snprintf(buf, cap, "%s", v.c_str());
  return E2D_OK;
}

extern "C" int
e2d_config_set_string(e2d_config * c, const char * section, const char * name, const char * value)
{
  if (!c || !section || !name || !value)
    return E2D_ERR_INVALID;
  c->cfg.kv[make_key(section, name)] = value;
  return E2D_OK;
}

extern "C" int
e2d_params_setup(e2d_params * out, const e2d_config * c)
{
  if (!out || !c)
    return E2D_ERR_INVALID;
  setup(c->cfg, out);
  return E2D_OK;
}

extern "C" int
e2d_params_from_ini(const char * path, e2d_params * out)
{
  if (!path || !out)
    return E2D_ERR_INVALID;
  Config cfg;
  FILE * f = std::fopen(path, "r");
  if (f)
  {
    parse_ini([f](char * buf, int n) { return std::fgets(buf, n, f) != nullptr; }, cfg.kv);
    std::fclose(f);
  }
  setup(cfg, out);
  return f ? E2D_OK : E2D_ERR_IO;
}

extern "C" int
e2d_params_from_string(const char * ini_text, e2d_params * out)
{
  if (!ini_text || !out)
    return E2D_ERR_INVALID;
  Config       cfg;
  const char * cur = ini_text;
  // emulate fgets: up to n-1 characters, stopping after a newline
  parse_ini(
    [&cur](char * buf, int n) {
      if (!*cur)
        return false;
      int k = 0;
      while (k < n - 1 && *cur)
      {
        buf[k++] = *cur;
        if (*cur++ == '\n')
          break;
      }
      buf[k] = '\0';
      return true;
    },
    cfg.kv);
  setup(cfg, out);
  return E2D_OK;
}

extern "C" int
e2d_params_print(const e2d_params * p)
{
  if (!p)
    return E2D_ERR_INVALID;
  std::printf("##########################\n");
  std::printf("Simulation run parameters:\n");
  std::printf("##########################\n");
  std::printf("nx         : %d\n", p->nx);
  std::printf("ny         : %d\n", p->ny);
  std::printf("dx         : %f\n", p->dx);
  std::printf("dy         : %f\n", p->dy);
  std::printf("imin       : %d\n", p->imin);
  std::printf("imax       : %d\n", p->imax);
  std::printf("jmin       : %d\n", p->jmin);
  std::printf("jmax       : %d\n", p->jmax);
  std::printf("nStepmax   : %d\n", p->nStepmax);
  std::printf("tEnd       : %f\n", p->tEnd);
  std::printf("nOutput    : %d\n", p->nOutput);
  std::printf("gamma0     : %f\n", p->gamma0);
  std::printf("cfl        : %f\n", p->cfl);
  std::printf("smallr     : %12.10f\n", p->smallr);
  std::printf("smallc     : %12.10f\n", p->smallc);
  std::printf("slope_type : %f\n", p->slope_type);
  std::printf("riemann    : %d\n", p->riemannSolverType);
  std::printf("problem    : %d\n", p->problemType);
  std::printf("implementation version : %d\n", p->implementationVersion);
  std::printf("##########################\n");
  return E2D_OK;
}
