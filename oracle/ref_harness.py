"""Live-reference harness (TEST INFRASTRUCTURE, dev container only).

Imports the genuine VolatileSpace physics modules from /root/reference so the
oracle restatement (oracle/vs_oracle.c) can be pinned against the reference's
own numba path and golden vectors can be dumped into tests/golden/.

/root/reference does not exist on the GPU box: nothing that runs there may
import this module.  Only oracle/gen_golden.py and the "live" (auto-skipped)
CPU tests use it.

Recipe follows SURVEY.md section 8(c):
  1. stub `pygame` (defaults.py:37-54 reads K_* constants, phys_body.py:162
     asks the display surface for its size),
  2. writable CWD holding settings.ini with numba=True / fastmath=False
     (peripherals.py:633-656, defaults.py:31-32),
  3. writable NUMBA_CACHE_DIR (reference tree is read-only, cache=True).
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("VSB_REFERENCE_ROOT", "/root/reference")
_WORKDIR = os.environ.get("VSB_REF_WORKDIR", "/tmp/vsb200_ref_work")

_loaded = None


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "volatilespace", "physics"))


def _install_pygame_stub():
    if "pygame" in sys.modules:
        return

    class _Anything(types.ModuleType):
        def __getattr__(self, name):
            if name.startswith("K_"):
                return abs(hash(name)) % 512
            raise AttributeError(name)

    pg = _Anything("pygame")

    class _Surface:
        @staticmethod
        def get_size():
            return (1366, 768)

    display = types.ModuleType("pygame.display")
    display.get_surface = lambda: _Surface()
    pg.display = display
    sys.modules["pygame"] = pg
    sys.modules["pygame.display"] = display


def load(curve_points=300):
    """Import the reference physics modules; returns a namespace of them."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError("reference tree not present at " + REFERENCE_ROOT)
    os.makedirs(_WORKDIR, exist_ok=True)
    os.makedirs(os.path.join(_WORKDIR, "numba_cache"), exist_ok=True)
    os.environ.setdefault("NUMBA_CACHE_DIR", os.path.join(_WORKDIR, "numba_cache"))
    with open(os.path.join(_WORKDIR, "settings.ini"), "w") as f:
        f.write("[game]\nnumba = True\nfastmath = False\npredict_coi_limit = 300\n"
                "\n[graphics]\ncurve_points = %d\n" % curve_points)
    _install_pygame_stub()
    old_cwd = os.getcwd()
    os.chdir(_WORKDIR)          # peripherals reads settings.ini relative to CWD
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        from volatilespace import peripherals, defaults
        from volatilespace.physics import (phys_editor, phys_body, phys_vessel,
                                           phys_shared, enhanced_kepler_solver)
    finally:
        os.chdir(old_cwd)
        sys.path.remove(REFERENCE_ROOT)
    ns = types.SimpleNamespace(
        peripherals=peripherals, defaults=defaults, phys_editor=phys_editor,
        phys_body=phys_body, phys_vessel=phys_vessel, phys_shared=phys_shared,
        enhanced_kepler_solver=enhanced_kepler_solver,
        solar_system_ini=os.path.join(REFERENCE_ROOT, "Resources", "BuiltinMaps",
                                      "Solar System.ini"))
    _loaded = ns
    return ns


def set_curve_points(ns, physics_obj, p):
    """Change P on a reference Physics object without touching settings.ini
    (phys_editor.py:153-157, phys_body.py:160-167)."""
    import numpy as np
    physics_obj.curve_points = int(p)
    if hasattr(physics_obj, "ell_t"):
        physics_obj.ell_t = np.linspace(-np.pi, np.pi, p)
        physics_obj.par_t = np.linspace(-np.pi - 1, np.pi + 1, p)
    else:
        physics_obj.t = np.linspace(-np.pi, np.pi, p)


def editor_from_arrays(ns, mass, den, pos, vel, conf=None):
    """Build a reference editor Physics from raw arrays (phys_editor.py:169-201)
    and run the trailing per-frame calls once so `types` exists."""
    import numpy as np
    n = len(mass)
    conf = dict(ns.defaults.sim_config) if conf is None else dict(conf)
    body_data = {
        "name": np.array(["b%d" % i for i in range(n)]),
        "mass": np.array(mass, dtype=np.float64),
        "den": np.array(den, dtype=np.float64),
        "color": np.zeros((n, 3), int),
        "atm_pres0": np.zeros(n), "atm_scale_h": np.zeros(n), "atm_den0": np.zeros(n),
    }
    body_orb = {"kepler": False, "pos": np.array(pos, dtype=np.float64),
                "vel": np.array(vel, dtype=np.float64)}
    ph = ns.phys_editor.Physics()
    ph.load_system(conf, body_data, body_orb)
    ph.kepler_basic()
    ph.temp_color()
    ph.classify()
    return ph


def editor_frame(ph, warp=1):
    """One faithful editor frame (editor.py:1503-1523)."""
    deleted = []
    for _ in range(warp):
        ph.gravity()
        ph.body()
        d = ph.inelastic_collision()
        if d is not None:
            deleted.append(int(d))
    ph.kepler_basic()
    ph.temp_color()
    ph.classify()
    return deleted
