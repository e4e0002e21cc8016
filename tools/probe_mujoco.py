"""Probe the GPU box's image for any MuJoCo binding (VERDICT r1 item 1a / SURVEY §8c).

Run under gpurun; writes gpurun_out/probe_mujoco.txt.  If anything imports, tools/dump_mujoco_golden.py is the next step.
"""
import importlib, os, subprocess, sys

out = []
def p(*a):
    s = " ".join(str(x) for x in a)
    out.append(s)
    print(s, flush=True)

p("python", sys.version.replace("\n", " "))
for name in ["mujoco_py", "mujoco", "dm_control", "gymnasium", "gym", "mujoco.mjx", "brax", "pybullet", "isaacgym",
             "qpsolvers", "daqp", "quadprog", "osqp", "cvxopt", "transforms3d", "scipy", "pysdf", "trimesh"]:
    try:
        m = importlib.import_module(name)
        p("IMPORT OK  ", name, getattr(m, "__version__", "?"), getattr(m, "__file__", "?"))
    except Exception as e:  # noqa: BLE001
        p("IMPORT FAIL", name, type(e).__name__, str(e)[:120])

def sh(cmd, timeout=120):
    try:
        r = subprocess.run(cmd, shell=True, capture_output=True, text=True, timeout=timeout)
        p("$", cmd, "-> rc", r.returncode)
        for line in (r.stdout + r.stderr).strip().splitlines()[:40]:
            p("   ", line[:200])
    except Exception as e:  # noqa: BLE001
        p("$", cmd, "-> EXC", e)

sh("find / -xdev \\( -iname '*mujoco*' -o -iname 'libmujoco*' -o -iname 'mjkey*' -o -iname '*libccd*' \\) "
   "-not -path '/proc/*' -not -path '*/gpurun_out/*' 2>/dev/null | head -40", 300)
sh("python -m pip download --no-deps -d /tmp/mjdl mujoco 2>&1 | tail -3", 60)
sh("python -m pip download --no-deps -d /tmp/mjdl --no-index --find-links /opt/wheelhouse mujoco 2>&1 | tail -3", 60)
sh("ls /opt/wheelhouse 2>/dev/null | grep -i -E 'mujoco|gym|dm_control|brax|qpsolvers|daqp' | head")
sh("python -m pip list 2>/dev/null | grep -i -E 'mujoco|gym|dm.control|brax|bullet|qpsolvers|daqp|quadprog|osqp' | head")
sh("ls ~/.mujoco /root/.mujoco /opt/mujoco* /usr/local/mujoco* 2>&1 | head")
os.makedirs("gpurun_out", exist_ok=True)
open("gpurun_out/probe_mujoco.txt", "w").write("\n".join(out) + "\n")
