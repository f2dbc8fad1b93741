"""The reference's input decks (src/test_*.ini, src/sedov_blast/test_sedov_blast_2d.ini) as data.

Values are the parity fixtures of BASELINE.json's configs; they are rendered to .ini text on demand
(``deck_text``) and parsed by the same strtof-based reader as any user file, so no deck files need
to travel with the package.  Override any key with ``section__key=value``.
"""
from __future__ import annotations

_HYDRO = dict(gamma0="1.666", cfl="0.8", niter_riemann="10", slope_type="2", riemann="hllc")


def _deck(problem, prefix, run, mesh, hydro=None, extra=None):
    d = {"run": dict(run), "mesh": dict(mesh), "hydro": dict(_HYDRO, problem=problem, **(hydro or {})),
         "output": {"outputPrefix": prefix}, "other": {"implementationVersion": "0"}}
    for k, v in (extra or {}).items():
        d[k] = dict(v)
    return d


_REFLECT = dict(boundary_type_xmin="1", boundary_type_xmax="1", boundary_type_ymin="1", boundary_type_ymax="1")
_ABSORB = dict(boundary_type_xmin="2", boundary_type_xmax="2", boundary_type_ymin="2", boundary_type_ymax="2")

DECKS = {
    "implode": _deck("implode", "test_implode", dict(tEnd="10.0", nStepmax="100", nOutput="10"),
                     dict(nx="256", ny="128", xmin="-1.0", xmax="1.0", ymin="0.0", ymax="1.0", **_REFLECT)),
    "implode_big": _deck("implode", "test_implode_big", dict(tEnd="10.0", nStepmax="100", nOutput="-1"),
                         dict(nx="2048", ny="2048", **_REFLECT)),
    "blast": _deck("blast", "test_blast", dict(tEnd="10.0", nStepmax="500", nOutput="10"),
                   dict(nx="128", ny="192", xmin="0.0", xmax="1.0", ymin="0.0", ymax="1.5", **_REFLECT),
                   extra={"blast": dict(density_in="1.0", density_out="1.2")}),
    "blast_large": _deck("blast", "test_blast_large", dict(tEnd="10.0", nStepmax="10", nOutput="-1"),
                         dict(nx="1280", ny="1920", xmin="0.0", xmax="1.0", ymin="0.0", ymax="1.5", **_REFLECT),
                         extra={"blast": dict(density_in="1.0", density_out="1.2")}),
    "four_quadrant": _deck("four_quadrant", "test_four_quadrant", dict(tEnd="0.6", nStepmax="1200", nOutput="100"),
                           dict(nx="256", ny="256", xmin="0.0", xmax="1.0", ymin="0.0", ymax="1.0", **_ABSORB)),
    "discontinuity": _deck("discontinuity", "test_discontinuity", dict(tEnd="0.6", nStepmax="1200", nOutput="100"),
                           dict(nx="256", ny="256", xmin="0.0", xmax="1.0", ymin="0.0", ymax="1.0", **_ABSORB)),
    "shocked_bubble": _deck("shocked_bubble", "test_shocked_bubble", dict(tEnd="10.0", nStepmax="500", nOutput="10"),
                            dict(nx="445", ny="89", xmin="0.0", xmax="0.445", ymin="0.0", ymax="0.089",
                                 boundary_type_xmin="2", boundary_type_xmax="2", boundary_type_ymin="1",
                                 boundary_type_ymax="1"),
                            hydro=dict(gamma0="1.2", cfl="0.5")),
    "sedov_blast_2d": _deck("blast", "sedov_blast_2d", dict(tEnd="0.5", nStepmax="3000", nOutput="1000"),
                            dict(nx="256", ny="256", xmin="0.0", xmax="1.0", ymin="0.0", ymax="1.0", **_ABSORB),
                            hydro=dict(gamma0="1.4", iorder="2"),
                            extra={"blast": dict(density_in="1.0", density_out="1.0", pressure_in="1",
                                                 pressure_out="1e-7", radius="0.01", total_energy_inside="0.311357",
                                                 compute_radial_profile="yes", num_radial_bins="200")}),
}


def deck_text(name: str, **overrides) -> str:
    """Render deck ``name`` as .ini text; ``overrides`` are ``section__key=value`` (e.g. mesh__nx=8192)."""
    d = {s: dict(kv) for s, kv in DECKS[name].items()}
    for k, v in overrides.items():
        sec, key = k.split("__", 1)
        d.setdefault(sec, {})[key] = str(v)
    out = []
    for sec, kv in d.items():
        out.append(f"[{sec}]")
        out += [f"{k}={v}" for k, v in kv.items()]
        out.append("")
    return "\n".join(out)


def write_deck(path: str, name: str, **overrides) -> str:
    with open(path, "w") as f:
        f.write(deck_text(name, **overrides))
    return path
