#!/bin/bash
# Round-2 probe: is the reference's third-party dynamics (pybullet 3.2.7 / PyFlyt 0.11.1 /
# gymnasium / stable_baselines3 / h5py) present or installable on the GPU box?  Output -> profiles/r2_ref_probe.log
exec > gpurun_out/r2_ref_probe.log 2>&1
set -x
date -u
python --version
for m in pybullet PyFlyt gymnasium stable_baselines3 h5py numba optuna pynput pybullet_data; do
  python -c "import $m; print('$m', getattr($m,'__version__','?'), $m.__file__)" 2>&1 | tail -1
done
python -m pip --version
timeout 60 python -m pip download --no-deps -d /tmp/dl pybullet==3.2.7 2>&1 | tail -3
timeout 60 python -m pip install --no-index --find-links /opt/wheelhouse pybullet PyFlyt gymnasium 2>&1 | tail -3
ls /opt/wheelhouse 2>/dev/null | grep -i -E "bullet|flyt|gymnasium|stable|h5py" || echo "wheelhouse: none of pybullet/PyFlyt/gymnasium/stable_baselines3/h5py"
find / -xdev \( -iname "*pybullet*" -o -iname "*pyflyt*" -o -iname "cf2x*" -o -iname "*gymnasium*" -o -iname "libBullet*" -o -iname "bullet3*" \) -not -path "/proc/*" 2>/dev/null | head -20
echo "find done"
timeout 20 python - <<'PY'
import socket
try:
    socket.create_connection(("pypi.org", 443), timeout=5); print("network: pypi reachable")
except Exception as e:
    print("network: unreachable:", e)
PY
nproc; lscpu | head -20; free -g | head -2
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv
