"""
OpenSCENARIO ingest -- ``import_scenario`` with the reference's signature and conventions
(reference scenario_gym/xosc_interface/read.py:20-282, catalogs.py:30-84) on top of
``xml.etree.ElementTree``.  A JSON road network next to the file is loaded into
``scenario.road_network`` (surfaces only, see road_network.py); a missing file gives ``None`` as in
the reference (read.py:75-85).
"""
from __future__ import annotations

import os
import warnings
import xml.etree.ElementTree as ET
from functools import lru_cache
from typing import Dict, List, Optional, Tuple

import numpy as np

from .entity import ENTITY_CLASS_BY_TAG, BoundingBox, CatalogEntry, Entity, Pedestrian, Vehicle
from .road_network import RoadNetwork
from .scenario import Scenario
from .trajectory import Trajectory


def _properties(element) -> Tuple[dict, list]:
    props, files = {}, []
    prop = element.find("Properties")
    if prop is not None:
        for child in prop.findall("Property"):
            v = child.attrib["value"]
            try:
                v = float(v)
            except ValueError:
                pass
            props[child.attrib["name"]] = v
        for f in prop.findall("File"):
            files.append(f.attrib["filepath"])
    return props, files


def _load_object(element, catalog_name: Optional[str]) -> Optional[Entity]:
    """Catalog element -> entity with that catalog entry (reference catalogs.py:13-27)."""
    bb = element.find("BoundingBox")
    if bb is None:
        return None
    dims, center = bb.find("Dimensions"), bb.find("Center")
    box = BoundingBox(float(dims.attrib["width"]), float(dims.attrib["length"]),
                      float(center.attrib["x"]), float(center.attrib["y"]))
    cname = element.tag.lower() + "Category"
    props, files = _properties(element)
    entry = CatalogEntry(catalog_name, element.attrib["name"], element.attrib.get(cname),
                         element.tag, box, props, files)
    return ENTITY_CLASS_BY_TAG.get(element.tag, Entity)(entry)


_CATALOG_CACHE: Dict[Tuple[str, float, int], Tuple[str, Dict[str, Entity]]] = {}


def read_catalog_cached(catalog_file: str) -> Tuple[str, Dict[str, Entity]]:
    """
    ``read_catalog`` memoised on (path, mtime, size).  The reference re-parses every catalog
    file for every scenario (read.py:97-118), which is most of its ~17 ms per file; entries
    are copied before use (``ent.copy()`` in import_scenario), so sharing them is safe.
    """
    st = os.stat(catalog_file)
    key = (os.path.abspath(catalog_file), st.st_mtime, st.st_size)
    hit = _CATALOG_CACHE.get(key)
    if hit is None:
        hit = _CATALOG_CACHE[key] = read_catalog(catalog_file)
    return hit


def read_catalog(catalog_file: str) -> Tuple[str, Dict[str, Entity]]:
    """Catalog name and a dict entry name -> entity (reference catalogs.py:30-84)."""
    root = ET.parse(catalog_file).getroot()
    cat = root.find("Catalog")
    entries = {}
    for element in list(cat):
        ent = _load_object(element, cat.attrib["name"])
        if ent is not None:
            entries[ent.catalog_entry.catalog_entry] = ent
    return cat.attrib["name"], entries


def _traj_point(t: float, wp) -> np.ndarray:
    g = wp.attrib.get
    return np.array([t, float(wp.attrib["x"]), float(wp.attrib["y"]), float(g("z", np.nan)),
                     float(g("h", np.nan)), float(g("p", np.nan)), float(g("r", np.nan))])


def relabel_scenario(scenario: Scenario) -> Scenario:
    """ego, vehicle_i, pedestrian_i, other_i (reference read.py:244-272)."""
    vehicles = pedestrians = others = 0
    scenario.entities[0].ref = "ego"
    for e in scenario.entities[1:]:
        scenario._ref_to_entity.pop(e.ref, None)
        if isinstance(e, Vehicle):
            e.ref = f"vehicle_{vehicles}"
            vehicles += 1
        elif isinstance(e, Pedestrian):
            e.ref = f"pedestrian_{pedestrians}"
            pedestrians += 1
        else:
            e.ref = f"other_{others}"
            others += 1
        scenario._ref_to_entity[e.ref] = e
    scenario._ref_to_entity["ego"] = scenario.entities[0]
    return scenario


@lru_cache(maxsize=15)
def _road_network_from_json(filepath: str, mtime: float) -> RoadNetwork:
    return RoadNetwork.create_from_json(filepath)


def _load_road_network(root, cwd: str) -> Optional[RoadNetwork]:
    """RoadNetwork/SceneGraphFile or LogicFile of the scenario (reference read.py:65-85); JSON only."""
    node = root.find("RoadNetwork/SceneGraphFile")
    if node is None:
        node = root.find("RoadNetwork/LogicFile")
    if node is None:
        return None
    rn_path = node.attrib["filepath"]
    filepath = rn_path if os.path.isabs(rn_path) else os.path.join(cwd, rn_path)
    if os.path.splitext(filepath)[1] == "":
        filepath = f"{filepath}.json"
    if not filepath.endswith(".json") or not os.path.exists(filepath):
        return None
    return _road_network_from_json(filepath, os.path.getmtime(filepath))


def import_scenario(osc_file: str, relabel: bool = True, entity_types=None) -> Scenario:
    """Import a scenario from an OpenSCENARIO file (reference read.py:20-189)."""
    if not os.path.exists(osc_file):
        raise FileNotFoundError
    cwd = os.path.dirname(osc_file)
    root = ET.parse(osc_file).getroot()
    catalogs: Dict[str, Dict[str, Entity]] = {}
    for loc in root.iterfind("CatalogLocations/"):
        rel = loc.find("Directory").attrib["path"]
        path = rel if os.path.isabs(rel) else os.path.join(cwd, rel)
        if not os.path.isdir(path):
            continue
        for fn in os.listdir(path):
            if fn.endswith(".xosc"):
                name, entries = read_catalog_cached(os.path.join(path, fn))
                catalogs[name] = entries
    road_network = _load_road_network(root, cwd)
    entities: Dict[str, Entity] = {}
    for obj in root.iterfind("Entities/ScenarioObject"):
        ref = obj.attrib["name"]
        cat_ref = obj.find("CatalogReference")
        if cat_ref is None:
            ent = None
            for element in list(obj):
                ent = _load_object(element, None) or ent
            if ent is None:
                warnings.warn(f"Could not find a catalog reference or entry for entity {ref}.")
                continue
            ent.ref = ref
            entities[ref] = ent
        else:
            cname, ename = cat_ref.attrib["catalogName"], cat_ref.attrib["entryName"]
            if cname not in catalogs:
                warnings.warn(f"Could not find catalog: {cname}")
            elif ename not in catalogs[cname]:
                warnings.warn(f"Could not find entry {ename} in catalog {cname}.")
            else:
                ent = catalogs[cname][ename].copy()
                ent.ref = ref
                entities[ref] = ent
    for private in root.iterfind("Storyboard/Init/Actions/Private"):
        ref = private.attrib["entityRef"]
        for wp in private.iterfind("PrivateAction/TeleportAction/Position/WorldPosition"):
            if ref in entities:
                entities[ref].trajectory = Trajectory(np.stack([_traj_point(0, wp)], axis=0))
    for group in root.iterfind("Storyboard/Story/Act/ManeuverGroup"):
        eref = group.find("Actors/EntityRef")
        assert eref is not None, "Could not find entity reference in maneuver group."
        entity = entities.get(eref.attrib["entityRef"])
        if entity is None:
            continue
        for event in group.findall("Maneuver/Event"):
            action = event.find("Action/PrivateAction/RoutingAction/FollowTrajectoryAction")
            if action is None:
                continue
            vertices = action.findall("TrajectoryRef/Trajectory/Shape/Polyline/Vertex")
            vertices.extend(action.findall("Trajectory/Shape/Polyline/Vertex"))
            if vertices:
                pts = [_traj_point(float(v.attrib["time"]), v.find("Position/WorldPosition"))
                       for v in vertices]
                entity.trajectory = Trajectory(np.stack(pts, axis=0))
    header = root.find("FileHeader")
    props = {}
    if header is not None:
        props, files = _properties(header)
        if files and "files" not in props:
            props["files"] = files
    scenario = Scenario(list(entities.values()), name=os.path.splitext(os.path.basename(osc_file))[0],
                        road_network=road_network, properties=props)
    return relabel_scenario(scenario) if relabel else scenario


def _import_one(args) -> Scenario:
    path, relabel, entity_types = args
    return import_scenario(path, relabel=relabel, entity_types=entity_types)


def import_scenarios(osc_files, relabel: bool = True, entity_types=None, workers: Optional[int] = None):
    """
    Batched ingest: import many OpenSCENARIO files, catalogs parsed
    once per process, files spread over ``workers`` processes (default: serial for few or small
    files, else one per core; ``Scenario`` objects are picklable, as in the reference's ``mp.Pool``
    fan-out, tests/test_scenario_gym.py:152-160).  Order is preserved; a missing file raises
    ``FileNotFoundError`` like ``import_scenario``.
    """
    files = list(osc_files)
    for f in files:
        if not os.path.exists(f):
            raise FileNotFoundError(f)
    if workers is None:  # a pool only pays for many sizeable files (fork + pickling ~1 ms per scenario)
        big = len(files) >= 64 and sum(os.path.getsize(f) for f in files) >= 32768 * len(files)
        workers = min(os.cpu_count() or 1, 16) if big else 1
    if workers <= 1 or len(files) <= 1:
        return [import_scenario(f, relabel=relabel, entity_types=entity_types) for f in files]
    import multiprocessing as mp

    ctx = mp.get_context("fork")
    chunk = max(1, len(files) // (4 * workers))
    with ctx.Pool(workers) as pool:
        return pool.map(_import_one, [(f, relabel, entity_types) for f in files], chunksize=chunk)
