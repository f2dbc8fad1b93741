import sys
sys.path.insert(0, ".")
import euler2d_kokkos_b200 as e2d
from euler2d_kokkos_b200.decks import deck_text
for mode in ("strict", "fast"):
    hp = e2d.HydroParams.from_string(deck_text("implode", mesh__nx=300, mesh__ny=70, run__nOutput=-1, run__nStepmax=4, other__arithmetic=mode))
    with e2d.HydroRun(hp) as h:
        st = h.run()
        print(mode, st.nStep, st.t)
