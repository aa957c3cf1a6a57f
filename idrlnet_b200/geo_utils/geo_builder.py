"""Name -> shape factory (idrlnet/geo_utils/geo_builder.py:23-58)."""
from . import geo_obj
from .geo import Geometry

__all__ = ["GeometryBuilder"]


class GeometryBuilder:
    GEOMAP = {name: getattr(geo_obj, cls) for name, cls in {
        "Line1D": "Line1D", "Line": "Line", "Rectangle": "Rectangle", "Circle": "Circle", "Channel2D": "Tube2D",
        "Plane": "Plane", "Sphere": "Sphere", "Box": "Box", "Channel": "Tube3D", "Channel3D": "Tube3D",
        "Cylinder": "Cylinder", "CircularTube": "CircularTube", "Triangle": "Triangle", "Heart": "Heart"}.items()}

    @staticmethod
    def get_geometry(geo: str, **kwargs) -> Geometry:
        assert geo in GeometryBuilder.GEOMAP, f"The geometry {geo} not implemented!"
        return GeometryBuilder.GEOMAP[geo](**kwargs)
