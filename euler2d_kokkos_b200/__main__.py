"""python -m euler2d_kokkos_b200 <file.ini> — same command line as the reference's `euler2d` (src/main.cpp:65-71)."""
import sys

from .hydro_run import main_loop

if len(sys.argv) != 2:
    sys.stderr.write("Error: wrong number of argument; input filename must be the only parameter on the command line\n")
    sys.exit(1)
main_loop(sys.argv[1])
