#!/usr/bin/env python
"""Regenerates the committed golden fixtures in tests/golden/ from the reference tree.

Run in the build container only (needs /root/reference and oracle/_ref):
    python tests/golden/make_golden.py

What it writes
  integration_<cell>.json : inputs + expected nodal displacements (RESULT DESCRIPTION, tol 1e-12)
        of the reference's own ctest input files
          tests/input_files/solid_ele_hex8_Standard_eas_none_stressout.4C.yaml
          tests/input_files/solid_ele_hex27_Standard_stressout.4C.yaml
          tests/input_files/solid_ele_tet10_Standard_stressout.4C.yaml
  integration_<cell>_mulf.json : MULF prestress inputs tests/input_files/solid_ele_<cell>_Standard_mulf.4C.yaml
        (CoupNeoHooke + CoupAnisoExpo element fibre, PRESTRESSTIME 1) with their nodal displacements
  integration_hex8_eas_full.json / _eas_mild.json : TECH eas_full / eas_mild hex8 inputs
        tests/input_files/solid_ele_hex8_Standard_eas_{full_stressout_cauchy_strainout_ea,mild_stressout_2pk_strainout_gl}.4C.yaml
        (same fibre material) with nodal displacements, stresses and strains
  integration_hex8_loadlin.json : tests/input_files/loadlin.4C.yaml, 300 hex8 EAS-full elements under a
        follower pressure (TYPE orthopressure with LOADLIN); dispy of node 404 (1e-8)
  integration_hex8_isoogden.json : tests/input_files/elasthyper_isoogden.4C.yaml, MAT_ElastHyper{ELAST_IsoOgden,
        ELAST_VolOgden} (principal-stretch formulation) under prescribed displacements; nodal Cauchy stresses
  integration_hex8_druckerprager_RetToCone[_elastic_tang].json : tests/input_files/mat_druckerprager_RetToCone*.4C.yaml,
        40 hex8 KINEM linear elements with the history material MAT_Struct_DruckerPrager under nodal forces
  integration_cooks_fbar.json : Cook's membrane with 64 F-bar hex8 elements
          tests/input_files/sohex8fbar_cooks_nl_new_struc.4C.yaml (dispy of node 162, 1e-8)
  unit_known_answers.json : known answers of unittests/mat/4C_elasthyper_service_test.cpp:36-155,
        unittests/mat/4C_stvenantkirchhoff_test.cpp:44-118 and the glibc srand(0)/rand() input
        sequence of unittests/solid_ele/4C_solid_ele_calc_lib_test.cpp:243-343
  ref_elements.npz : element matrices / force vectors and material tangents computed HERE by
        oracle/_ref (the reference's own tensor, einsum and shape-function headers) on seeded
        random inputs, for all cell types x materials.
"""
import ctypes
import json
import os
import re
import sys

import numpy as np
import yaml

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)

CASES = {
    "hex8": "tests/input_files/solid_ele_hex8_Standard_eas_none_stressout.4C.yaml",
    "hex27": "tests/input_files/solid_ele_hex27_Standard_stressout.4C.yaml",
    "tet10": "tests/input_files/solid_ele_tet10_Standard_stressout.4C.yaml",
    "tet4": "tests/input_files/solid_ele_tet4_Standard_stressout.4C.yaml",
    "hex20": "tests/input_files/solid_ele_hex20_Standard_stressout.4C.yaml",
    "wedge6": "tests/input_files/solid_ele_wedge6_Standard_stressout.4C.yaml",
    "pyramid5": "tests/input_files/solid_ele_pyramid5_Standard_stressout.4C.yaml",
    "hex18": "tests/input_files/solid_ele_hex18_Standard_volume_neumann.4C.yaml",  # (no stressout input for hex18)
}


# KINEM linear + MAT_Struct_StVenantKirchhoff: the reference's own small-strain ctest inputs
LINEAR_CELLS = ("hex8", "hex20", "hex27", "tet4", "pyramid5")


def integration_case(cell, rel, linear=False):
    with open(os.path.join(REF, rel)) as fh:
        y = yaml.safe_load(fh)
    nodes = {}
    for line in y["NODE COORDS"]:
        m = re.match(r"NODE\s+(\d+)\s+COORD\s+(\S+)\s+(\S+)\s+(\S+)", line)
        nodes[int(m.group(1))] = [float(m.group(i)) for i in (2, 3, 4)]
    nn = max(nodes)
    coords = [nodes[i + 1] for i in range(nn)]
    eles = []
    for line in y["STRUCTURE ELEMENTS"]:
        tok = line.split()
        assert tok[1] == "SOLID" and tok[2].lower() == cell
        k = 3
        conn = []
        while tok[k] != "MAT":
            conn.append(int(tok[k]) - 1)
            k += 1
        assert ("KINEM nonlinear" in line) != linear and ("KINEM linear" in line) == linear
        eles.append(conn)
    surf = {1: [], 2: []}
    for line in y["DSURF-NODE TOPOLOGY"]:
        m = re.match(r"NODE\s+(\d+)\s+DSURFACE\s+(\d+)", line)
        surf[int(m.group(2))].append(int(m.group(1)) - 1)
    mats = {m["MAT"]: m for m in y["MATERIALS"]}
    if "MAT_Struct_StVenantKirchhoff" in mats[1]:  # the *_Standard_linear inputs
        matkind, nh = "MAT_Struct_StVenantKirchhoff", mats[1]["MAT_Struct_StVenantKirchhoff"]
    else:
        eh = mats[1]["MAT_ElastHyper"]
        assert eh["NUMMAT"] == 1
        matkind, nh = "ELAST_CoupNeoHooke", mats[eh["MATIDS"][0]]["ELAST_CoupNeoHooke"]
    dbc = y["DESIGN SURF DIRICH CONDITIONS"][0]
    neu = y["DESIGN SURF NEUMANN CONDITIONS"][0]
    assert dbc["E"] == 1 and neu["E"] == 2
    sd = y["STRUCTURAL DYNAMIC"]
    expected = []
    for r in y["RESULT DESCRIPTION"]:
        s = r["STRUCTURE"]
        if s["QUANTITY"] in ("dispx", "dispy", "dispz"):
            expected.append({"node": s["NODE"] - 1, "dof": "xyz".index(s["QUANTITY"][-1]),
                             "value": float(s["VALUE"]), "tol": float(s["TOLERANCE"])})
    # nodal stresses (IO: STRUCT_STRESS Cauchy, Gauss-point values extrapolated to the nodes)
    comp = {"xx": 0, "yy": 1, "zz": 2, "xy": 3, "yz": 4, "xz": 5}
    expected_stress = []
    for r in y["RESULT DESCRIPTION"]:
        s = r["STRUCTURE"]
        if s["QUANTITY"].startswith("stress_"):
            expected_stress.append({"node": s["NODE"] - 1, "comp": comp[s["QUANTITY"][-2:]],
                                    "value": float(s["VALUE"]), "tol": float(s["TOLERANCE"])})
    extra = {}
    if "DESIGN VOL NEUMANN CONDITIONS" in y:  # body load on every element (DVOL 1 = all nodes)
        vn = y["DESIGN VOL NEUMANN CONDITIONS"][0]
        assert vn["E"] == 1 and len(y["DVOL-NODE TOPOLOGY"]) == nn
        extra["vol_neumann"] = {"ONOFF": vn["ONOFF"], "VAL": vn["VAL"], "FUNCT": vn["FUNCT"]}
    return {
        **extra,
        "source": rel,
        "celltype": cell,
        "iostress": y.get("IO", {}).get("STRUCT_STRESS"),
        "expected_stress": expected_stress,
        "coords": coords,
        "elements": eles,
        "material": {"kind": matkind, "YOUNG": float(nh["YOUNG"]), "NUE": float(nh["NUE"])},
        "kinem": "linear" if linear else "nonlinear",
        "dirichlet_nodes": sorted(surf[1]),
        "neumann_nodes": sorted(surf[2]),
        "neumann": {"ONOFF": neu["ONOFF"], "VAL": neu["VAL"], "FUNCT": neu["FUNCT"],
                    "FUNCT1": y["FUNCT1"][0]["SYMBOLIC_FUNCTION_OF_SPACE_TIME"],
                    "FUNCT2": y["FUNCT2"][0]["SYMBOLIC_FUNCTION_OF_SPACE_TIME"]},
        "timestep": float(sd["TIMESTEP"]), "numstep": int(sd["NUMSTEP"]),
        "tolres": float(sd["TOLRES"]), "toldisp": float(sd["TOLDISP"]),
        "expected_disp": expected,
    }


def mulf_case(cell, rel, tech=None):
    """MULF prestress ctest input: one element block with MAT_ElastHyper{CoupNeoHooke, CoupAnisoExpo},
    element fibre FIBER1, PRESTRESS MULF until PRESTRESSTIME; known answers = nodal displacements."""
    with open(os.path.join(REF, rel)) as fh:
        y = yaml.safe_load(fh)
    nodes = {}
    for line in y["NODE COORDS"]:
        m = re.match(r"NODE\s+(\d+)\s+COORD\s+(\S+)\s+(\S+)\s+(\S+)", line)
        nodes[int(m.group(1))] = [float(m.group(i)) for i in (2, 3, 4)]
    coords = [nodes[i + 1] for i in range(max(nodes))]
    eles, fibers = [], []
    for line in y["STRUCTURE ELEMENTS"]:
        tok = line.split()
        assert tok[1] == "SOLID" and tok[2].lower() == cell
        assert (f"TECH {tech}" in line) if tech else ("PRESTRESS_TECH mulf" in line)
        k = 3
        conn = []
        while tok[k] != "MAT":
            conn.append(int(tok[k]) - 1)
            k += 1
        eles.append(conn)
        f = tok.index("FIBER1")
        fibers.append([float(tok[f + 1]), float(tok[f + 2]), float(tok[f + 3])])
    surf = {1: [], 2: []}
    for line in y["DSURF-NODE TOPOLOGY"]:
        m = re.match(r"NODE\s+(\d+)\s+DSURFACE\s+(\d+)", line)
        surf[int(m.group(2))].append(int(m.group(1)) - 1)
    mats = {m["MAT"]: m for m in y["MATERIALS"]}
    eh = mats[1]["MAT_ElastHyper"]
    summands = []
    for mid in eh["MATIDS"]:
        (name, par), = [(k, v) for k, v in mats[mid].items() if k != "MAT"]
        summands.append({"name": name, "params": par})
    dbc = y["DESIGN SURF DIRICH CONDITIONS"][0]
    neu = y["DESIGN SURF NEUMANN CONDITIONS"][0]
    assert dbc["E"] == 1 and neu["E"] == 2
    sd = y["STRUCTURAL DYNAMIC"]
    assert tech or sd["PRESTRESS"].lower() == "mulf"
    vol = None
    if "DESIGN VOL NEUMANN CONDITIONS" in y:
        vn = y["DESIGN VOL NEUMANN CONDITIONS"][0]
        vol_nodes = sorted(int(re.match(r"NODE\s+(\d+)\s+DVOL\s+(\d+)", l).group(1)) - 1
                           for l in y["DVOL-NODE TOPOLOGY"]
                           if int(re.match(r"NODE\s+(\d+)\s+DVOL\s+(\d+)", l).group(2)) == vn["E"])
        assert vol_nodes == list(range(len(coords)))  # the volume load acts on every element
        vol = {"ONOFF": vn["ONOFF"], "VAL": vn["VAL"], "FUNCT": vn["FUNCT"]}
    comp = {"xx": 0, "yy": 1, "zz": 2, "xy": 3, "yz": 4, "xz": 5}
    expected_stress, expected_strain = [], []
    for r in y["RESULT DESCRIPTION"]:
        s = r["STRUCTURE"]
        for key, lst in (("stress_", expected_stress), ("strain_", expected_strain)):
            if s["QUANTITY"].startswith(key):
                lst.append({"node": s["NODE"] - 1, "comp": comp[s["QUANTITY"][-2:]],
                            "value": float(s["VALUE"]), "tol": float(s["TOLERANCE"])})
    expected = []
    for r in y["RESULT DESCRIPTION"]:
        s = r["STRUCTURE"]
        if s["QUANTITY"] in ("dispx", "dispy", "dispz"):
            expected.append({"node": s["NODE"] - 1, "dof": "xyz".index(s["QUANTITY"][-1]),
                             "value": float(s["VALUE"]), "tol": float(s["TOLERANCE"])})
    return {
        "source": rel, "celltype": cell, "coords": coords, "elements": eles, "fibers": fibers,
        "summands": summands, "dirichlet_nodes": sorted(surf[1]), "neumann_nodes": sorted(surf[2]),
        "neumann": {"ONOFF": neu["ONOFF"], "VAL": neu["VAL"], "FUNCT": neu["FUNCT"],
                    "FUNCT1": y["FUNCT1"][0]["SYMBOLIC_FUNCTION_OF_SPACE_TIME"],
                    "FUNCT2": y["FUNCT2"][0]["SYMBOLIC_FUNCTION_OF_SPACE_TIME"]},
        "timestep": float(sd["TIMESTEP"]), "numstep": int(sd["NUMSTEP"]),
        "prestresstime": float(sd.get("PRESTRESSTIME", -1.0)), "tech": tech, "vol_neumann": vol,
        "iostress": y["IO"].get("STRUCT_STRESS"), "iostrain": y["IO"].get("STRUCT_STRAIN"),
        "tolres": float(sd["TOLRES"]), "toldisp": float(sd["TOLDISP"]),
        "expected_disp": expected, "expected_stress": expected_stress, "expected_strain": expected_strain,
    }


def loadlin_case(rel):
    """tests/input_files/loadlin.4C.yaml: cantilever of 300 hex8 EAS-full elements (CoupNeoHooke) under a
    follower pressure (DESIGN SURF NEUMANN CONDITIONS TYPE orthopressure, LOADLIN true), statics;
    known answer = dispy of node 404."""
    with open(os.path.join(REF, rel)) as fh:
        y = yaml.safe_load(fh)
    nodes = {}
    for line in y["NODE COORDS"]:
        m = re.match(r"NODE\s+(\d+)\s+COORD\s+(\S+)\s+(\S+)\s+(\S+)", line)
        nodes[int(m.group(1))] = [float(m.group(i)) for i in (2, 3, 4)]
    coords = [nodes[i + 1] for i in range(max(nodes))]
    eles = []
    for line in y["STRUCTURE ELEMENTS"]:
        tok = line.split()
        assert tok[1] == "SOLID" and tok[2] == "HEX8" and "KINEM nonlinear TECH eas_full" in line
        eles.append([int(t) - 1 for t in tok[3:11]])
    dbc, = y["DESIGN SURF DIRICH CONDITIONS"]
    neu, = y["DESIGN SURF NEUMANN CONDITIONS"]
    assert neu["TYPE"] == "orthopressure" and dbc["ONOFF"] == [1, 1, 1]
    surf = {dbc["E"]: [], neu["E"]: []}
    for line in y["DSURF-NODE TOPOLOGY"]:
        m = re.match(r"NODE\s+(\d+)\s+DSURFACE\s+(\d+)", line)
        if int(m.group(2)) in surf:
            surf[int(m.group(2))].append(int(m.group(1)) - 1)
    mats = {m["MAT"]: m for m in y["MATERIALS"]}
    eh = mats[1]["MAT_ElastHyper"]
    nh = mats[eh["MATIDS"][0]]["ELAST_CoupNeoHooke"]
    var, = [f for f in y["FUNCT1"] if "VARIABLE" in f]
    assert var["TYPE"] == "linearinterpolation"
    sd = y["STRUCTURAL DYNAMIC"]
    assert sd["DYNAMICTYPE"] == "Statics" and sd["LOADLIN"] is True
    expected = []
    for r in y["RESULT DESCRIPTION"]:
        s = r["STRUCTURE"]
        expected.append({"node": s["NODE"] - 1, "dof": "xyz".index(s["QUANTITY"][-1]),
                         "value": float(s["VALUE"]), "tol": float(s["TOLERANCE"])})
    return {
        "source": rel, "celltype": "hex8", "tech": "eas_full", "coords": coords, "elements": eles,
        "material": {"kind": "ELAST_CoupNeoHooke", "YOUNG": float(nh["YOUNG"]), "NUE": float(nh["NUE"])},
        "dirichlet_nodes": sorted(surf[dbc["E"]]), "neumann_nodes": sorted(surf[neu["E"]]),
        "neumann": {"TYPE": neu["TYPE"], "ONOFF": neu["ONOFF"][:3], "VAL": neu["VAL"][:3],
                    "FUNCT": neu["FUNCT"][:3],
                    "FUNCT1_TIMES": [float(v) for v in var["TIMES"]],
                    "FUNCT1_VALUES": [float(v) for v in var["VALUES"]]},
        "timestep": float(sd["TIMESTEP"]), "numstep": int(sd["NUMSTEP"]),
        "tolres": float(sd["TOLRES"]), "toldisp": float(sd["TOLDISP"]),
        "expected_disp": expected,
    }


def isoogden_case(rel):
    """tests/input_files/elasthyper_isoogden.4C.yaml: three cubes of 8 hex8 each, MAT_ElastHyper{IsoOgden, VolOgden},
    driven by prescribed displacements (compression, extension, shear); known answers = nodal Cauchy stresses."""
    with open(os.path.join(REF, rel)) as fh:
        y = yaml.safe_load(fh)
    ids, coords = [], []
    for line in y["NODE COORDS"]:
        m = re.match(r"NODE\s+(\d+)\s+COORD\s+(\S+)\s+(\S+)\s+(\S+)", line)
        ids.append(int(m.group(1)))
        coords.append([float(m.group(i)) for i in (2, 3, 4)])
    index = {g: k for k, g in enumerate(ids)}
    eles = []
    for line in y["STRUCTURE ELEMENTS"]:
        tok = line.split()
        assert tok[1] == "SOLID" and tok[2] == "HEX8" and line.rstrip('"').endswith("KINEM nonlinear")
        eles.append([index[int(t)] for t in tok[3:11]])
    surf = {}
    for line in y["DSURF-NODE TOPOLOGY"]:
        m = re.match(r"NODE\s+(\d+)\s+DSURFACE\s+(\d+)", line)
        surf.setdefault(int(m.group(2)), []).append(index[int(m.group(1))])
    dirichlet = [{"nodes": sorted(surf[c["E"]]), "ONOFF": c["ONOFF"], "VAL": c["VAL"], "FUNCT": c["FUNCT"]}
                 for c in y["DESIGN SURF DIRICH CONDITIONS"]]
    functs = {}
    for k in (1, 2, 3):
        var, = [f for f in y[f"FUNCT{k}"] if "VARIABLE" in f]
        assert var["TYPE"] == "linearinterpolation"
        functs[str(k)] = {"TIMES": [float(v) for v in var["TIMES"]], "VALUES": [float(v) for v in var["VALUES"]]}
    mats = {m["MAT"]: m for m in y["MATERIALS"]}
    summands = []
    for mid in mats[1]["MAT_ElastHyper"]["MATIDS"]:
        (name, par), = [(k, v) for k, v in mats[mid].items() if k != "MAT"]
        summands.append({"name": name, "params": par})
    comp = {"xx": 0, "yy": 1, "zz": 2, "xy": 3, "yz": 4, "xz": 5}
    expected = [{"node": index[r["STRUCTURE"]["NODE"]], "comp": comp[r["STRUCTURE"]["QUANTITY"][-2:]],
                 "value": float(r["STRUCTURE"]["VALUE"]), "tol": float(r["STRUCTURE"]["TOLERANCE"])}
                for r in y["RESULT DESCRIPTION"]]
    sd = y["STRUCTURAL DYNAMIC"]
    return {"source": rel, "celltype": "hex8", "coords": coords, "elements": eles, "summands": summands,
            "dirichlet": dirichlet, "functs": functs, "iostress": y["IO"]["STRUCT_STRESS"],
            "timestep": float(sd["TIMESTEP"]), "numstep": int(sd["NUMSTEP"]),
            "tolres": float(sd["TOLRES"]), "toldisp": float(sd["TOLDISP"]), "expected_stress": expected}


def druckerprager_case(rel):
    """tests/input_files/mat_druckerprager_RetToCone[_elastic_tang].4C.yaml: cantilever of 40 hex8 KINEM linear
    elements, MAT_Struct_DruckerPrager (history material), nodal forces (DESIGN POINT NEUMANN) growing with t,
    statics; known answers = tip displacements."""
    with open(os.path.join(REF, rel)) as fh:
        y = yaml.safe_load(fh)
    ids, coords = [], []
    for line in y["NODE COORDS"]:
        m = re.match(r"NODE\s+(\d+)\s+COORD\s+(\S+)\s+(\S+)\s+(\S+)", line)
        ids.append(int(m.group(1)))
        coords.append([float(m.group(i)) for i in (2, 3, 4)])
    index = {g: k for k, g in enumerate(ids)}
    eles = []
    for line in y["STRUCTURE ELEMENTS"]:
        tok = line.split()
        assert tok[1] == "SOLID" and tok[2] == "HEX8" and "KINEM linear" in line
        eles.append([index[int(t)] for t in tok[3:11]])
    dbc, = y["DESIGN SURF DIRICH CONDITIONS"]
    pn, = y["DESIGN POINT NEUMANN CONDITIONS"]
    assert dbc["ONOFF"] == [1, 1, 1] and dbc["VAL"] == [0, 0, 0]
    dnodes = [index[int(re.match(r"NODE\s+(\d+)\s+DNODE\s+(\d+)", l).group(1))] for l in y["DNODE-NODE TOPOLOGY"]
              if int(re.match(r"NODE\s+(\d+)\s+DNODE\s+(\d+)", l).group(2)) == pn["E"]]
    snodes = [index[int(re.match(r"NODE\s+(\d+)\s+DSURFACE\s+(\d+)", l).group(1))] for l in y["DSURF-NODE TOPOLOGY"]
              if int(re.match(r"NODE\s+(\d+)\s+DSURFACE\s+(\d+)", l).group(2)) == dbc["E"]]
    assert y["FUNCT1"][0]["SYMBOLIC_FUNCTION_OF_TIME"] == "t"
    mat = dict(y["MATERIALS"][0]["MAT_Struct_DruckerPrager"])
    mat = {k: (float(v) if k != "TANG" else v) for k, v in mat.items()}
    sd = y["STRUCTURAL DYNAMIC"]
    expected = [{"node": index[r["STRUCTURE"]["NODE"]], "dof": "xyz".index(r["STRUCTURE"]["QUANTITY"][-1]),
                 "value": float(r["STRUCTURE"]["VALUE"]), "tol": float(r["STRUCTURE"]["TOLERANCE"])}
                for r in y["RESULT DESCRIPTION"]]
    return {"source": rel, "celltype": "hex8", "kinem": "linear", "coords": coords, "elements": eles, "material": mat,
            "dirichlet_nodes": sorted(snodes),
            "point_neumann": {"nodes": sorted(dnodes), "ONOFF": pn["ONOFF"], "VAL": pn["VAL"], "FUNCT": pn["FUNCT"]},
            "timestep": float(sd["TIMESTEP"]), "numstep": int(sd["NUMSTEP"]),
            "tolres": float(sd["TOLRES"]), "toldisp": float(sd["TOLDISP"]), "expected_disp": expected}


def necking_case(rel):
    """tests/input_files/plastic_necking_eas.4C.yaml: a quarter of a round bar of 96 EAS-mild hex8 elements,
    MAT_Struct_PlasticNlnLogNeoHooke (finite-strain von Mises plasticity, a history material), pulled by a prescribed
    end displacement u_z = t in 30 steps; symmetry conditions on points / lines / surfaces.  Lower-dimensional Dirichlet
    entities override higher ones for all dofs of their nodes (Core::FE::Utils::Dbc hierarchy), so the list is stored
    surfaces first, points last and applied in that order."""
    with open(os.path.join(REF, rel)) as fh:
        y = yaml.safe_load(fh)
    ids, coords = [], []
    for line in y["NODE COORDS"]:
        m = re.match(r"NODE\s+(\d+)\s+COORD\s+(\S+)\s+(\S+)\s+(\S+)", line)
        ids.append(int(m.group(1)))
        coords.append([float(m.group(i)) for i in (2, 3, 4)])
    index = {g: k for k, g in enumerate(ids)}
    eles = []
    for line in y["STRUCTURE ELEMENTS"]:
        tok = line.split()
        assert tok[1] == "SOLID" and tok[2] == "HEX8" and "KINEM nonlinear" in line and "TECH eas_mild" in line
        eles.append([index[int(t)] for t in tok[3:11]])
    dirichlet = []
    for sec, topo, key in (("DESIGN SURF DIRICH CONDITIONS", "DSURF-NODE TOPOLOGY", "DSURFACE"),
                           ("DESIGN LINE DIRICH CONDITIONS", "DLINE-NODE TOPOLOGY", "DLINE"),
                           ("DESIGN POINT DIRICH CONDITIONS", "DNODE-NODE TOPOLOGY", "DNODE")):
        for c in y[sec]:
            nodes = [index[int(re.match(rf"NODE\s+(\d+)\s+{key}\s+(\d+)", l).group(1))] for l in y[topo]
                     if int(re.match(rf"NODE\s+(\d+)\s+{key}\s+(\d+)", l).group(2)) == c["E"]]
            dirichlet.append({"nodes": sorted(nodes), "ONOFF": c["ONOFF"], "VAL": c["VAL"], "FUNCT": c["FUNCT"],
                              "level": ("DSURFACE", "DLINE", "DNODE").index(key)})
    assert y["FUNCT1"][0]["SYMBOLIC_FUNCTION_OF_SPACE_TIME"] == "t"
    mat = {k: float(v) for k, v in y["MATERIALS"][0]["MAT_Struct_PlasticNlnLogNeoHooke"].items()}
    assert "HARDENING_FUNC" not in mat  # parameter hardening only
    sd = y["STRUCTURAL DYNAMIC"]
    expected = [{"node": index[r["STRUCTURE"]["NODE"]], "dof": "xyz".index(r["STRUCTURE"]["QUANTITY"][-1]),
                 "value": float(r["STRUCTURE"]["VALUE"]), "tol": float(r["STRUCTURE"]["TOLERANCE"])}
                for r in y["RESULT DESCRIPTION"] if r["STRUCTURE"]["QUANTITY"] in ("dispx", "dispy", "dispz")]
    return {"source": rel, "celltype": "hex8", "tech": "eas_mild", "coords": coords, "elements": eles, "material": mat,
            "dirichlet": dirichlet, "functs": {"1": {"TIMES": [0.0, 1000.0], "VALUES": [0.0, 1000.0]}},
            "timestep": float(sd["TIMESTEP"]), "numstep": int(sd["NUMSTEP"]),
            "tolres": float(sd["TOLRES"]), "toldisp": float(sd["TOLDISP"]), "expected_disp": expected}


def visco_creep_case(rel):
    """tests/input_files/viscoNeoHooke_creep.4C.yaml: unit cube of 64 hex8 EAS-mild elements, MAT_VISCONEOHOOKE (finite-strain
    history material), constant surface traction, symmetry Dirichlet conditions per dof, default time integration of
    STRUCTURAL DYNAMIC (GenAlpha, rho_inf = 1, no damping, consistent mass with DENS 1e-10), 100 steps of 0.1."""
    with open(os.path.join(REF, rel)) as fh:
        y = yaml.safe_load(fh)
    ids, coords = [], []
    for line in y["NODE COORDS"]:
        m = re.match(r"NODE\s+(\d+)\s+COORD\s+(\S+)\s+(\S+)\s+(\S+)", line)
        ids.append(int(m.group(1)))
        coords.append([float(m.group(i)) for i in (2, 3, 4)])
    index = {g: k for k, g in enumerate(ids)}
    eles = []
    for line in y["STRUCTURE ELEMENTS"]:
        tok = line.split()
        assert tok[1] == "SOLID" and tok[2] == "HEX8" and "KINEM nonlinear TECH eas_mild" in line and "MAT 1" in line
        eles.append([index[int(t)] for t in tok[3:11]])
    topo = {"DNODE": {}, "DLINE": {}, "DSURFACE": {}}
    for key, kind in (("DNODE-NODE TOPOLOGY", "DNODE"), ("DLINE-NODE TOPOLOGY", "DLINE"), ("DSURF-NODE TOPOLOGY", "DSURFACE")):
        for line in y.get(key, []):
            m = re.match(r"NODE\s+(\d+)\s+" + kind + r"\s+(\d+)", line)
            topo[kind].setdefault(int(m.group(2)), []).append(index[int(m.group(1))])
    fixed = set()
    for key, kind in (("DESIGN POINT DIRICH CONDITIONS", "DNODE"), ("DESIGN LINE DIRICH CONDITIONS", "DLINE"),
                      ("DESIGN SURF DIRICH CONDITIONS", "DSURFACE")):
        for c in y.get(key, []):
            assert all(v == 0 for v in c["VAL"])
            for n in topo[kind][c["E"]]:
                for d in range(3):
                    if c["ONOFF"][d]:
                        fixed.add(3 * n + d)
    neu, = y["DESIGN SURF NEUMANN CONDITIONS"]
    assert all(f == 0 for f in neu["FUNCT"])
    mat = {k: float(v) for k, v in y["MATERIALS"][0]["MAT_VISCONEOHOOKE"].items()}
    sd = y["STRUCTURAL DYNAMIC"]
    assert "DYNAMICTYPE" not in sd and "DAMPING" not in sd and "GENALPHA" not in y   # the defaults named above
    expected = [{"node": index[r["STRUCTURE"]["NODE"]], "dof": "xyz".index(r["STRUCTURE"]["QUANTITY"][-1]),
                 "value": float(r["STRUCTURE"]["VALUE"]), "tol": float(r["STRUCTURE"]["TOLERANCE"])}
                for r in y["RESULT DESCRIPTION"]]
    return {"source": rel, "celltype": "hex8", "tech": "eas_mild", "coords": coords, "elements": eles, "material": mat,
            "dirichlet_nodes": [], "dirichlet_dofs": sorted(fixed),
            "neumann_nodes": sorted(topo["DSURFACE"][neu["E"]]),
            "neumann": {"ONOFF": neu["ONOFF"][:3], "VAL": neu["VAL"][:3], "FUNCT": neu["FUNCT"][:3], "FUNCT1": "t", "FUNCT2": None},
            "timestep": float(sd["TIMESTEP"]), "numstep": int(sd["NUMSTEP"]), "toldisp": float(sd["TOLDISP"]),
            "genalpha": {"rho_inf": 1.0, "genavg": "TrLike", "damping": "None", "predict": "ConstDis", "lumpmass": False},
            "expected_disp": expected}


def cooks_case(rel, tech, kinem="nonlinear"):
    """Cook's membrane inputs (64 hex8, CoupNeoHooke nu=0.4999): several Dirichlet conditions with
    per-dof ONOFF, one dead surface load, known answer = dispy of node 162 (and internal energy)."""
    with open(os.path.join(REF, rel)) as fh:
        y = yaml.safe_load(fh)
    nodes = {}
    for line in y["NODE COORDS"]:
        m = re.match(r"NODE\s+(\d+)\s+COORD\s+(\S+)\s+(\S+)\s+(\S+)", line)
        nodes[int(m.group(1))] = [float(m.group(i)) for i in (2, 3, 4)]
    coords = [nodes[i + 1] for i in range(max(nodes))]
    eles = []
    for line in y["STRUCTURE ELEMENTS"]:
        tok = line.split()
        assert tok[1] == "SOLID" and tok[2] == "HEX8" and f"TECH {tech}" in line and (("KINEM " + kinem) in line)
        eles.append([int(t) - 1 for t in tok[3:11]])
    surf = {}
    for line in y["DSURF-NODE TOPOLOGY"]:
        m = re.match(r"NODE\s+(\d+)\s+DSURFACE\s+(\d+)", line)
        surf.setdefault(int(m.group(2)), []).append(int(m.group(1)) - 1)
    mats = {m["MAT"]: m for m in y["MATERIALS"]}
    if "MAT_ElastHyper" in mats[1]:
        eh = mats[1]["MAT_ElastHyper"]
        assert eh["NUMMAT"] == 1
        nh = mats[eh["MATIDS"][0]]["ELAST_CoupNeoHooke"]
        matkind = "ELAST_CoupNeoHooke"
    else:
        nh = mats[1]["MAT_Struct_StVenantKirchhoff"]
        matkind = "MAT_Struct_StVenantKirchhoff"
    fixed = set()
    for dbc in y["DESIGN SURF DIRICH CONDITIONS"]:
        assert all(v == 0 for v in dbc["VAL"])
        for n in surf[dbc["E"]]:
            for d in range(3):
                if dbc["ONOFF"][d]:
                    fixed.add(3 * n + d)
    (neu,) = y["DESIGN SURF NEUMANN CONDITIONS"]
    sd = y["STRUCTURAL DYNAMIC"]
    expected, energy = [], None
    for r in y["RESULT DESCRIPTION"]:
        s = r["STRUCTURE"]
        if s["QUANTITY"] in ("dispx", "dispy", "dispz"):
            expected.append({"node": s["NODE"] - 1, "dof": "xyz".index(s["QUANTITY"][-1]),
                             "value": float(s["VALUE"]), "tol": float(s["TOLERANCE"])})
        if s["QUANTITY"] == "internal_energy":
            energy = {"value": float(s["VALUE"]), "tol": float(s["TOLERANCE"])}
    return {
        "source": rel, "celltype": "hex8", "tech": tech, "kinem": kinem, "coords": coords, "elements": eles,
        "material": {"kind": matkind, "YOUNG": float(nh["YOUNG"]), "NUE": float(nh["NUE"])},
        "dirichlet_dofs": sorted(fixed), "dirichlet_nodes": [],
        "neumann_nodes": sorted(surf[neu["E"]]),
        "neumann": {"ONOFF": neu["ONOFF"] + [0] * 3, "VAL": neu["VAL"] + [0] * 3,
                    "FUNCT": neu["FUNCT"] + [0] * 3,
                    "FUNCT1": y["FUNCT1"][0]["SYMBOLIC_FUNCTION_OF_SPACE_TIME"], "FUNCT2": None},
        "timestep": float(sd["TIMESTEP"]), "numstep": int(sd["NUMSTEP"]),
        "tolres": float(sd["TOLRES"]), "toldisp": float(sd["TOLDISP"]),
        "normcombi": sd.get("NORMCOMBI_RESFDISP", "And"),
        "expected_disp": expected, "expected_internal_energy": energy,
    }


def unit_known_answers():
    libc = ctypes.CDLL("libc.so.6")
    libc.srand(0)
    RAND_MAX = 2147483647
    seq = [libc.rand() / RAND_MAX for _ in range(2 * 3 * 27)]
    return {
        "gamma_delta": {
            "source": "unittests/mat/4C_elasthyper_service_test.cpp:36-81",
            "prinv": [3.1, 3.2, 1.05], "dPI": [0.536, 0.681, 3.981],
            "ddPII": [0.153, 0.861, 2.964, 1.549, 0.682, 3.584],
            "gamma_ref": [5.294200000000001, -1.362, 8.360099999999999],
            "delta_ref": [125.31604, -25.0124, 23.03238, 3.444, -6.5058, 29.79144, -16.7202, -2.724],
            "tol": 1e-10},
        "right_cauchy_green": {
            "source": "unittests/mat/4C_elasthyper_service_test.cpp:83-106",
            "E": [0.1, 0.2, 0.3, 0.01, 0.02, 0.03], "C_ref": [1.2, 1.4, 1.6, 0.02, 0.04, 0.06],
            "tol": 1e-10},
        "isoneohooke_invariant_derivatives": {
            "source": "unittests/mat/4C_elasthyper_service_test.cpp:108-155",
            "MUE": 1.3, "prinv": [3.1, 3.2, 1.05], "dPI_ref": [0.639514, 0.0, -0.629363],
            "ddPII_ref": [0.0, 0.0, 0.799191, 0.0, -0.20302, 0.0], "tol": 1e-4},
        "stvenantkirchhoff": {
            "source": "unittests/mat/4C_stvenantkirchhoff_test.cpp:44-118",
            "YOUNG": 210.0, "NUE": 0.3, "E": [1.0, 1.0, 1.0, 0.5, 0.5, 0.5],
            "strain_energy_ref": 908.6538, "tol": 1e-4},
        "bfree_stiffness": {
            "source": "unittests/solid_ele/4C_solid_ele_calc_lib_test.cpp:243-343",
            "glibc_srand0_rand_over_RAND_MAX": seq,
            "xi": [0.1, 0.2, 0.3], "pk2_full": [[1.1, 1.2, 1.3], [1.2, 2.2, 2.3], [1.3, 2.3, 3.3]],
            "tol": 1e-13},
    }


def ref_elements():
    from oracle import oracle as orc
    assert orc.ref_lib() is not None, "oracle/_ref not built"
    rng = np.random.default_rng(20261017)
    out = {}
    mats = {orc.MAT_COUP_NEOHOOKE: [200.0, 0.25], orc.MAT_ISO_VOL: [1.0, 1000.0],
            orc.MAT_SVK: [1000.0, 0.3]}
    # (new cells are appended so that the random stream of the earlier fixtures stays what it was)
    for ct, name in ((orc.HEX8, "hex8"), (orc.HEX27, "hex27"), (orc.TET10, "tet10"), (orc.TET4, "tet4"), (orc.HEX20, "hex20"),
                     (orc.WEDGE6, "wedge6"), (orc.PYRAMID5, "pyramid5"), (orc.HEX18, "hex18")):
        n = orc.num_nodes(ct)
        X = orc.node_param_coords(ct) * np.array([1.0, 1.1, 1.2]) + 0.05 * rng.uniform(-1, 1, (n, 3))
        u = 0.04 * rng.uniform(-1, 1, (n, 3))
        out[f"{name}_X"] = X
        out[f"{name}_u"] = u
        for mk, par in mats.items():
            Ke, fe, err = orc.ref_element(ct, mk, par, X, u)
            assert err == 0
            out[f"{name}_m{mk}_K"] = Ke
            out[f"{name}_m{mk}_f"] = fe
            out[f"m{mk}_par"] = np.array(par)
    # material tangents
    E = np.array([0.11, -0.07, 0.05, 0.02, -0.03, 0.04])
    out["mat_E"] = E
    import ctypes as C
    dp = C.POINTER(C.c_double)
    for mk, par in mats.items():
        S = np.zeros(6)
        Cm = np.zeros(36)
        p = np.array(par)
        orc.ref_lib().ref_material(mk, p.ctypes.data_as(dp), E.ctypes.data_as(dp),
                                   S.ctypes.data_as(dp), Cm.ctypes.data_as(dp))
        out[f"mat{mk}_S"] = S
        out[f"mat{mk}_C"] = Cm.reshape(6, 6).T.copy()
    # the unit-test cmat built by the reference's einsum_sym
    S = np.array([1.1, 2.2, 3.3, 1.2, 2.3, 1.3])
    Cm = np.zeros(36)
    orc.ref_lib().ref_test_cmat_from_pk2(S.ctypes.data_as(dp), Cm.ctypes.data_as(dp))
    out["unit_pk2"] = S
    out["unit_cmat"] = Cm.reshape(6, 6).T.copy()
    return out


if __name__ == "__main__":
    for cell, rel in CASES.items():
        with open(os.path.join(HERE, f"integration_{cell}.json"), "w") as fh:
            json.dump(integration_case(cell, rel), fh, indent=1)
    for cell in LINEAR_CELLS:
        with open(os.path.join(HERE, f"integration_{cell}_linear.json"), "w") as fh:
            json.dump(integration_case(cell, f"tests/input_files/solid_ele_{cell}_Standard_linear.4C.yaml", linear=True), fh, indent=1)
    for name, rel, tech in (
            ("cooks_fbar", "tests/input_files/sohex8fbar_cooks_nl_new_struc.4C.yaml", "fbar"),):
        with open(os.path.join(HERE, f"integration_{name}.json"), "w") as fh:
            json.dump(cooks_case(rel, tech), fh)
    for tech in ("shell_ans", "shell_eas", "shell_eas_ans"):
        with open(os.path.join(HERE, f"integration_hex8_{tech}.json"), "w") as fh:
            json.dump(integration_case("hex8", f"tests/input_files/solid_ele_hex8_Standard_{tech}.4C.yaml"), fh, indent=1)
    for tech, name in (("eas_mild", "sohex8_easmild_cooks_lin"), ("eas_full", "sohex8_easfull_cooks_lin")):
        with open(os.path.join(HERE, f"integration_cooks_{tech}_linear.json"), "w") as fh:
            json.dump(cooks_case(f"tests/input_files/{name}.4C.yaml", tech, "linear"), fh)
    for cell in ("hex8", "hex27", "tet10", "tet4", "hex20", "wedge6", "pyramid5"):
        with open(os.path.join(HERE, f"integration_{cell}_mulf.json"), "w") as fh:
            json.dump(mulf_case(cell, f"tests/input_files/solid_ele_{cell}_Standard_mulf.4C.yaml"), fh, indent=1)
    for tech, rel in (("eas_full", "tests/input_files/solid_ele_hex8_Standard_eas_full_stressout_cauchy_strainout_ea.4C.yaml"),
                      ("eas_mild", "tests/input_files/solid_ele_hex8_Standard_eas_mild_stressout_2pk_strainout_gl.4C.yaml")):
        with open(os.path.join(HERE, f"integration_hex8_{tech}.json"), "w") as fh:
            json.dump(mulf_case("hex8", rel, tech), fh, indent=1)
    for name, cell in (("pyramid5_fbar", "pyramid5"), ("hex8_fbar", "hex8"), ("hex8_mulf_fbar", "hex8")):
        with open(os.path.join(HERE, f"integration_{name}.json"), "w") as fh:
            json.dump(mulf_case(cell, f"tests/input_files/solid_ele_{cell}_Standard_{name[len(cell) + 1:]}.4C.yaml", "fbar"), fh, indent=1)
    with open(os.path.join(HERE, "integration_hex8_loadlin.json"), "w") as fh:
        json.dump(loadlin_case("tests/input_files/loadlin.4C.yaml"), fh)
    with open(os.path.join(HERE, "integration_hex8_isoogden.json"), "w") as fh:
        json.dump(isoogden_case("tests/input_files/elasthyper_isoogden.4C.yaml"), fh)
    for name in ("RetToCone", "RetToCone_elastic_tang"):
        with open(os.path.join(HERE, f"integration_hex8_druckerprager_{name}.json"), "w") as fh:
            json.dump(druckerprager_case(f"tests/input_files/mat_druckerprager_{name}.4C.yaml"), fh)
    for name in ("plastic_necking_eas",):
        with open(os.path.join(HERE, f"integration_hex8_{name}.json"), "w") as fh:
            json.dump(necking_case(f"tests/input_files/{name}.4C.yaml"), fh)
    with open(os.path.join(HERE, "integration_hex8_visco_creep.json"), "w") as fh:
        json.dump(visco_creep_case("tests/input_files/viscoNeoHooke_creep.4C.yaml"), fh)
    with open(os.path.join(HERE, "unit_known_answers.json"), "w") as fh:
        json.dump(unit_known_answers(), fh, indent=1)
    np.savez_compressed(os.path.join(HERE, "ref_elements.npz"), **ref_elements())
    print("golden fixtures written to", HERE)
