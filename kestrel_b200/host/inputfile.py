"""Reader for Kestrel's block-structured input files (Input.f90:54-565).

``Block:`` headers, ``key = value`` lines, ``#`` comments anywhere, ``%`` comments
at line start.  Keys are case-insensitive inside Domain / Parameters / Solver /
Output / Topog and case-sensitive for capX, cubeLength, sourceFlux ...
(Input.f90:232-275).  Unknown keys are skipped with a warning (Messages.f90:299).
Only what the time step needs is interpreted; georeferencing (Lat/Lon), raster
DEMs, NetCDF / KML options are parsed and ignored.
"""
from __future__ import annotations

import math
import warnings
from typing import List

from .settings import Cap, Cube, FluxSource, RunSet

_BLOCKS = ("Domain", "Source", "Cap", "Cube", "Parameters", "Solver", "Output", "Topog")


def _read_set(s: str) -> List[float]:  # "(a, b, c)"
    s = s.strip()
    if s.startswith("("):
        s = s[1:]
    if s.endswith(")"):
        s = s[:-1]
    return [float(x) for x in s.split(",") if x.strip()]


def _real(s: str) -> float:
    return float(s.strip().replace("d", "e").replace("D", "e"))


def read_input_file(path: str) -> RunSet:
    rs = RunSet()
    block = None
    blocks = {b: [] for b in ("Domain", "Parameters", "Solver", "Output", "Topog")}
    caps, cubes, srcs = [], [], []
    with open(path) as fh:
        for raw in fh:
            line = raw.strip()
            if "#" in line:
                line = line[: line.index("#")].strip()
            if not line or line[0] in "%#":
                continue
            if ":" in line:
                name = line[: line.index(":")].strip()
                if name in _BLOCKS:
                    block = name
                    if name == "Cap":
                        caps.append({})
                    elif name == "Cube":
                        cubes.append({})
                    elif name == "Source":
                        srcs.append({})
                continue
            if "=" not in line or block is None:
                continue
            key, val = line.split("=", 1)
            key, val = key.strip(), val.strip()
            if block == "Cap":
                caps[-1][key] = val
            elif block == "Cube":
                cubes[-1][key] = val
            elif block == "Source":
                srcs[-1][key] = val
            else:
                blocks[block].append((key.lower(), val))

    # ---- Domain (DomainSettings.f90:86-153)
    for k, v in blocks["Domain"]:
        if k == "nxtiles": rs.nXtiles = int(v)
        elif k == "nytiles": rs.nYtiles = int(v)
        elif k == "nxpertile": rs.nXpertile = int(v)
        elif k == "nypertile": rs.nYpertile = int(v)
        elif k == "xtilesize": rs.Xtilesize = _real(v)
        elif k == "ytilesize":
            rs.Ytilesize = _real(v)
            rs.Xtilesize = None
        elif k == "boundary conditions": rs.bcs = v.lower()
        elif k == "boundary hn": rs.bcsHnval = _real(v)
        elif k == "boundary u": rs.bcsuval = _real(v)
        elif k == "boundary v": rs.bcsvval = _real(v)
        elif k == "boundary psi": rs.bcspsival = _real(v)
        elif k in ("lat", "latitude", "lon", "longitude"): pass
        else: warnings.warn(f"Input label unrecognized: {k}")
    if rs.Xtilesize is None:
        rs.Xtilesize = rs.Ytilesize * float(rs.nXpertile) / float(rs.nYpertile)
        rs.Ytilesize = None if rs.Ytilesize is None else rs.Ytilesize

    # ---- Parameters (Parameters.f90:171-467)
    pmap = {
        "g": "g", "chezy co": "ChezyCo", "manning co": "ManningCo", "coulomb co": "CoulombCo",
        "pouliquen min": "PouliquenMinSlope", "pouliquen max": "PouliquenMaxSlope",
        "pouliquen intermediate": "PouliquenIntermediateSlope", "pouliquen beta": "PouliquenBeta",
        "edwards2019 betastar": "Edwards2019betastar", "edwards2019 kappa": "Edwards2019kappa",
        "edwards2019 gamma": "Edwards2019Gamma", "voellmy switch rate": "VoellmySwitchRate",
        "voellmy switch value": "VoellmySwitchValue", "erosion rate": "EroRate",
        "granular erosion rate": "EroRateGranular", "erosion depth": "EroDepth",
        "erosion critical height": "EroCriticalHeight", "bed porosity": "BedPorosity", "rhow": "rhow",
        "rhos": "rhos", "maxpack": "maxPack", "max pack": "maxPack", "solid diameter": "SolidDiameter",
        "eddy viscosity": "EddyViscosity", "settling speed": "ws0",
    }
    for k, v in blocks["Parameters"]:
        if k in pmap: setattr(rs, pmap[k], _real(v))
        elif k == "drag": rs.drag = v.lower()
        elif k == "erosion": rs.erosion = v.lower()
        elif k == "deposition": rs.deposition = v.lower()
        elif k == "erosion transition": rs.erosion_transition = v.lower()
        elif k == "morphodynamic damping": rs.morpho_damp = v.lower()
        elif k == "switch function": rs.fswitch = v.lower()
        elif k in ("iverson", "geometric factors"):
            if v.lower() == "off": rs.geometric_factors = False
            elif v.lower() == "on": rs.geometric_factors = True
        else: warnings.warn(f"Input label unrecognized: {k}")

    # ---- Solver (SolverSettings.f90:94-173)
    for k, v in blocks["Solver"]:
        if k == "t end": rs.tend = _real(v)
        elif k == "t start": rs.tstart = _real(v)
        elif k == "limiter": rs.limiter = v.lower()
        elif k == "height threshold": rs.heightThreshold = _real(v)
        elif k == "tile buffer": rs.TileBuffer = int(v)
        elif k == "cfl": rs.cfl = _real(v)
        elif k == "max dt": rs.maxdt = _real(v)
        elif k == "sponge strength": rs.SpongeStrength = _real(v)
        elif k in ("restart", "initial condition"): pass
        else: warnings.warn(f"Input label unrecognized: {k}")

    # ---- Output (OutputSettings.f90)
    for k, v in blocks["Output"]:
        if k == "n out": rs.Nout = int(v)
        elif k == "directory": rs.out_dir = v

    # ---- Topog (TopogSettings.f90:92-122)
    for k, v in blocks["Topog"]:
        if k == "type": rs.topog_type = v.lower()
        elif k == "topog function": rs.topog_func = v.lower()
        elif k == "topog params": rs.topog_params = _read_set(v)

    rs.finalize()

    # ---- Caps (InitConds.f90:254-420)
    for c in caps:
        cap = Cap()
        cap.x = _real(c.get("capX", "0")); cap.y = _real(c.get("capY", "0"))
        cap.u = _real(c.get("capU", "0")); cap.v = _real(c.get("capV", "0"))
        cap.psi = _real(c.get("capConc", "0"))
        shp = c.get("capShape", "para").lower()  # capShape_d = 'para', InitConds.f90:38
        cap.shape = {"parabolic": "para", "para": "para", "flat": "flat", "level": "level"}.get(shp, "para")
        hasR, hasH, hasV = "capRadius" in c, "capHeight" in c, "capVolume" in c
        if hasR: cap.radius = _real(c["capRadius"])
        if hasH: cap.height = _real(c["capHeight"])
        if hasV and not (hasR and hasH): cap.volume = _real(c["capVolume"])
        f = 1.0 if cap.shape == "flat" else 0.5
        if hasR and hasV and not hasH:
            cap.height = cap.volume / f / math.pi / cap.radius / cap.radius
        if hasH and hasV and not hasR:
            cap.radius = math.sqrt(cap.volume / f / math.pi / cap.height)
        rs.caps.append(cap)
    # ---- Cubes (InitConds.f90:526-773)
    for c in cubes:
        cu = Cube()
        cu.x = _real(c.get("cubeX", "0")); cu.y = _real(c.get("cubeY", "0"))
        cu.length = _real(c.get("cubeLength", "0")); cu.width = _real(c.get("cubeWidth", "0"))
        cu.height = _real(c.get("cubeHeight", "0"))
        cu.u = _real(c.get("cubeU", "0")); cu.v = _real(c.get("cubeV", "0"))
        cu.psi = _real(c.get("cubeConc", "0"))
        cu.shape = c.get("cubeShape", "flat").lower()
        rs.cubes.append(cu)
    # ---- Sources (InitConds.f90:46-200)
    for s in srcs:
        fs = FluxSource()
        fs.x = _real(s.get("sourceX", "0")); fs.y = _real(s.get("sourceY", "0"))
        fs.radius = _real(s["sourceRadius"])
        fs.time = _read_set(s["sourceTime"]); fs.flux = _read_set(s["sourceFlux"]); fs.psi = _read_set(s["sourceConc"])
        assert len(fs.time) == len(fs.flux) == len(fs.psi), "sourceTime/Flux/Conc sets differ in length"
        rs.sources.append(fs)
    return rs


def write_input_file(rs: RunSet, path: str, header: str = "") -> None:
    """Emit a RunSet in Kestrel's input format (normalised: one key per line, no comments
    other than the header).  read_input_file(write_input_file(rs)) round-trips."""
    def g(v):
        return repr(float(v))
    L = []
    if header:
        L += [f"% {ln}" for ln in header.splitlines()]
    L += ["Domain:", "Lat = 0", "Lon = 0", f"nXtiles = {rs.nXtiles}", f"nYtiles = {rs.nYtiles}",
          f"nXpertile = {rs.nXpertile}", f"nYpertile = {rs.nYpertile}", f"Xtilesize = {g(rs.Xtilesize)}",
          f"Boundary Conditions = {rs.bcs}"]
    if rs.bcs == "dirichlet":
        L += [f"Boundary Hn = {g(rs.bcsHnval)}", f"Boundary u = {g(rs.bcsuval)}", f"Boundary v = {g(rs.bcsvval)}",
              f"Boundary psi = {g(rs.bcspsival)}"]
    for c in rs.caps:
        L += ["", "Cap:", f"capX = {g(c.x)}", f"capY = {g(c.y)}", f"capRadius = {g(c.radius)}", f"capHeight = {g(c.height)}",
              f"capU = {g(c.u)}", f"capV = {g(c.v)}", f"capConc = {g(c.psi)}", f"capShape = {c.shape}"]
    for c in rs.cubes:
        L += ["", "Cube:", f"cubeX = {g(c.x)}", f"cubeY = {g(c.y)}", f"cubeLength = {g(c.length)}", f"cubeWidth = {g(c.width)}",
              f"cubeHeight = {g(c.height)}", f"cubeU = {g(c.u)}", f"cubeV = {g(c.v)}", f"cubeConc = {g(c.psi)}",
              f"cubeShape = {c.shape}"]
    for s in rs.sources:
        st = lambda xs: "(" + ", ".join(g(x) for x in xs) + ")"
        L += ["", "Source:", f"sourceX = {g(s.x)}", f"sourceY = {g(s.y)}", f"sourceRadius = {g(s.radius)}",
              f"sourceTime = {st(s.time)}", f"sourceFlux = {st(s.flux)}", f"sourceConc = {st(s.psi)}"]
    L += ["", "Parameters:", f"Drag = {rs.drag}", f"Erosion = {rs.erosion}", f"Deposition = {rs.deposition}",
          f"Erosion Transition = {rs.erosion_transition}", f"Morphodynamic damping = {rs.morpho_damp}",
          f"Switch function = {rs.fswitch}", f"Geometric factors = {'on' if rs.geometric_factors else 'off'}"]
    for key, attr in [("g", "g"), ("Chezy Co", "ChezyCo"), ("Manning Co", "ManningCo"), ("Coulomb Co", "CoulombCo"),
                      ("Pouliquen Min", "PouliquenMinSlope"), ("Pouliquen Max", "PouliquenMaxSlope"),
                      ("Pouliquen Intermediate", "PouliquenIntermediateSlope"), ("Pouliquen beta", "PouliquenBeta"),
                      ("Edwards2019 betastar", "Edwards2019betastar"), ("Edwards2019 kappa", "Edwards2019kappa"),
                      ("Edwards2019 gamma", "Edwards2019Gamma"), ("Voellmy switch rate", "VoellmySwitchRate"),
                      ("Voellmy switch value", "VoellmySwitchValue"), ("Erosion Rate", "EroRate"),
                      ("Granular Erosion Rate", "EroRateGranular"), ("Erosion depth", "EroDepth"),
                      ("Erosion critical height", "EroCriticalHeight"), ("Bed porosity", "BedPorosity"),
                      ("rhow", "rhow"), ("rhos", "rhos"), ("maxPack", "maxPack"), ("Solid diameter", "SolidDiameter"),
                      ("Eddy Viscosity", "EddyViscosity")]:
        L.append(f"{key} = {g(getattr(rs, attr))}")
    L += ["", "Solver:", f"T start = {g(rs.tstart)}", f"T end = {g(rs.tend)}", f"limiter = {rs.limiter}",
          f"Height threshold = {g(rs.heightThreshold)}", f"Tile Buffer = {rs.TileBuffer}", f"cfl = {g(rs.cfl)}"]
    if rs.maxdt < 1e300:
        L.append(f"max dt = {g(rs.maxdt)}")
    L += ["", "Output:", f"N out = {rs.Nout}", f"directory = {rs.out_dir}", "",
          "Topog:", "Type = Function", f"Topog function = {rs.topog_func}",
          "Topog params = (" + ", ".join(g(x) for x in rs.topog_params) + ")", ""]
    with open(path, "w") as fh:
        fh.write("\n".join(L))
