"""
TEST INFRASTRUCTURE ONLY -- import shim that lets the *unmodified* reference
package at /root/reference run in this container.

The reference imports five third-party roots that are absent here and cannot be
installed (no network): shapely, lxml, scenariogeneration, pyxodr, matplotlib.
``install()`` registers a ``sys.meta_path`` finder that fabricates them:

  * ``shapely``            -> ``oracle/refshim/mini_shapely.py`` (restated subset)
  * ``lxml.etree``         -> ``xml.etree.ElementTree`` with ``getchildren()``
  * ``scenariogeneration``, ``pyxodr``, ``matplotlib`` -> empty permissive stubs
    (only used by writers / plotting / xodr import, none on the rollout path)

Road-network files are not loaded (``RoadNetwork.create_from_file`` raises
``FileNotFoundError`` which ``xosc_interface/read.py:84-85`` suppresses), so
``scenario.road_network`` is ``None`` exactly as in SURVEY.md section 8c.

Nothing in the shipped product imports this module; it only exists so that
``oracle/gen_golden.py`` can execute the reference and commit golden vectors.
"""
from __future__ import annotations

import importlib.abc
import importlib.machinery
import os
import sys
import types
import xml.etree.ElementTree as _ET

REFERENCE_ROOT = os.environ.get("SCENARIO_GYM_REFERENCE", "/root/reference")

_STUB_ROOTS = ("shapely", "lxml", "scenariogeneration", "pyxodr", "matplotlib")


class _AnyMeta(type):
    def __getattr__(cls, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return type(name, (_Anything,), {})


class _Anything(metaclass=_AnyMeta):
    """Permissive placeholder: any attribute / call returns another placeholder."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Anything()

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Anything()


class _StubModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        val = type(name, (_Anything,), {})
        setattr(self, name, val)
        return val


class _Element(_ET.Element):
    """ElementTree element with lxml's ``getchildren``."""

    def getchildren(self):
        return list(self)


def _parse(source, parser=None):
    builder = _ET.TreeBuilder(element_factory=_Element)
    return _ET.parse(source, parser=_ET.XMLParser(target=builder))


def _make_lxml_etree() -> types.ModuleType:
    m = types.ModuleType("lxml.etree")
    m.Element = _Element
    m.parse = _parse
    m.fromstring = lambda s: _ET.fromstring(
        s, parser=_ET.XMLParser(target=_ET.TreeBuilder(element_factory=_Element))
    )
    m.tostring = _ET.tostring
    m.SubElement = _ET.SubElement
    m.ElementTree = _ET.ElementTree
    return m


def _make_shapely(fullname: str) -> types.ModuleType:
    from . import mini_shapely as ms

    m = _StubModule(fullname)
    exports = {
        "shapely": {},
        "shapely.geometry": dict(
            Point=ms.Point,
            Polygon=ms.Polygon,
            LineString=ms.LineString,
            LinearRing=ms.LinearRing,
            MultiPolygon=ms.MultiPolygon,
        ),
        "shapely.validation": dict(make_valid=ms.make_valid),
        "shapely.geometry.base": dict(BaseGeometry=ms.BaseGeometry),
        "shapely.strtree": dict(STRtree=ms.STRtree),
        "shapely.vectorized": dict(contains=ms.contains),
        "shapely.ops": dict(nearest_points=ms.nearest_points, unary_union=ms.unary_union),
    }
    for k, v in exports.get(fullname, {}).items():
        setattr(m, k, v)
    return m


class _Finder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path=None, target=None):
        if fullname.split(".")[0] in _STUB_ROOTS:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        name = spec.name
        if name == "lxml.etree":
            return _make_lxml_etree()
        if name.split(".")[0] == "shapely":
            return _make_shapely(name)
        return _StubModule(name)

    def exec_module(self, module):
        module.__path__ = []


_installed = False


def install() -> None:
    """Make ``import scenario_gym`` resolve to the reference, with stubs."""
    global _installed
    if _installed:
        return
    if not os.path.isdir(os.path.join(REFERENCE_ROOT, "scenario_gym")):
        raise RuntimeError(
            f"reference not found under {REFERENCE_ROOT}; the shim only works in "
            "the authoring container"
        )
    sys.meta_path.insert(0, _Finder())
    sys.path.insert(0, REFERENCE_ROOT)
    import scenario_gym.road_network as rn  # noqa: E402

    def _no_file(cls, filepath):
        raise FileNotFoundError(filepath)

    rn.RoadNetwork.create_from_file = classmethod(_no_file)
    _installed = True


class EmptyRoadNetwork:
    """Stand-in for ``RoadNetwork()`` with no geometry (areas are zero)."""

    def __init__(self):
        from .mini_shapely import MultiPolygon

        self.walkable_surface = MultiPolygon()
        self.impenetrable_surface = MultiPolygon()
        self.driveable_surface = MultiPolygon()
