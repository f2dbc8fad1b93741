"""Quick device-side timing of e2d_run on a deck (development aid; bench.py is the contract)."""
import sys
sys.path.insert(0, ".")
import euler2d_kokkos_b200 as e2d
from euler2d_kokkos_b200.decks import deck_text

deck = sys.argv[1] if len(sys.argv) > 1 else "four_quadrant"
nx = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
ny = int(sys.argv[3]) if len(sys.argv) > 3 else nx
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 20
mode = sys.argv[5] if len(sys.argv) > 5 else "strict"  # `[other] arithmetic`
hp = e2d.HydroParams.from_string(deck_text(deck, mesh__nx=nx, mesh__ny=ny, run__nOutput=-1, run__nStepmax=100000,
                                           run__tEnd=1e9, other__arithmetic=mode))
h = e2d.HydroRun(hp)
h.run(5)
for rep in range(3):
    st0 = h.run(5 + (rep + 1) * steps)
    print(f"{mode} {deck} {nx}x{ny}: {steps} steps in {st0.seconds*1e3:.2f} ms -> {steps*nx*ny/st0.seconds*1e-6:.1f} Mcell/s "
          f"({steps*nx*ny*64/st0.seconds*1e-9:.1f} GB/s algorithmic), launches={st0.launches}", flush=True)
