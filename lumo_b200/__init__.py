"""lumo_b200 — B200-native (sm_100a) implementation of lumo's rendering hot path.
Public names mirror the reference crate root (src/lib.rs:13-22, src/tracer.rs:1-14)."""
from .api import (Scene, Camera, CameraBuilder, CameraType, Material, Texture, Rectangle, Sphere, TriangleMesh, Face, Mesh,
                  Instance, LooseTriangles, Integrator, Renderer, SamplerType, ToneMap, PixelFilter, ColorSpace, illuminants)
from .spectrum import Spectrum
from .film import Film
from .image import Image
from . import parser
