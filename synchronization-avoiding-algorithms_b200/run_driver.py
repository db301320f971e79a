#!/usr/bin/env python
"""Run one of the reference's driver scripts UNCHANGED on top of this package.

    python  <pkg>/run_driver.py  /path/to/Data_prepare.py
    mpirun -np 2 python <pkg>/run_driver.py Data_prepare.py
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 <pkg>/run_driver.py Data_prepare.py

`python Data_prepare.py` would put the script's own directory — and with it the reference's `Tools/` — in front of
PYTHONPATH.  This launcher instead puts the package directory (the drop-in `Tools`) first and the `compat/`
stand-ins (mpi4py / meshio / h5py / mgmetis / matplotlib, used only where the real packages are not installed)
last, then executes the script as `__main__` in the current working directory (the scripts use relative paths:
Mesh_info/, Results/, Distributed_save/).
"""
import os
import runpy
import sys

PKG = os.path.dirname(os.path.abspath(__file__))


def main():
    if len(sys.argv) < 2:
        raise SystemExit(__doc__)
    script = sys.argv[1]
    sys.argv = sys.argv[1:]
    here = os.path.dirname(os.path.abspath(script))
    sys.path[:] = [p for p in sys.path if os.path.abspath(p or ".") not in (here, PKG)]
    sys.path.insert(0, PKG)
    sys.path.append(os.path.join(PKG, "compat"))
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()
