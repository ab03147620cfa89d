"""Film: host-side finalisation of the device accumulators (src/tracer/film.rs:173-210).

The device hands back `pixels[H,W,4]` = (sum r*w, sum g*w, sum b*w, sum w) and `splats[H,W,3]`;
`rgb_image` applies pixel value = sum rgb*w / sum w + splat_scale * splats / filter integral, then
the colour space's transfer function with truncating u8 conversion, exactly like the reference.
`save` writes an 8-bit RGB PNG (zlib only)."""
import struct
import zlib
import numpy as np
from . import color as _color


class Film:
    def __init__(self, pixels, splats, samples, pixel_filter, color_space, counters=None, stats=None):
        self.pixels, self.splats = pixels, splats
        self.resolution = (pixels.shape[1], pixels.shape[0])
        self.splat_scale = 1.0 / float(samples)
        self.filter, self.cs = pixel_filter, color_space
        self.counters = counters or {}
        self.stats = stats or {}

    def linear_rgb(self):
        with np.errstate(invalid="ignore", divide="ignore"):
            direct = self.pixels[..., :3] / self.pixels[..., 3:4]        # Pixel::value, film.rs:82-90
        splat = self.splat_scale * self.splats / self.filter.integral()  # film.rs:179-181
        return direct + splat

    def rgb_image(self):
        return _color.encode(self.linear_rgb(), self.cs)

    def rgb_image_device(self, ctx=None, device=0):
        """The same image computed by the CUDA film kernel (lumo_gpu_film_encode): byte for byte `rgb_image` — the transfer
        curve's pow is csrc/common/lumo_math.h on both sides."""
        from . import native
        own = ctx is None
        if own: ctx = native.GpuContext(device)
        try:
            return ctx.film_encode(self.pixels, self.splats, self.splat_scale, self.filter.integral(), 1 if self.cs == 2 else 0)
        finally:
            if own: ctx.close()

    def save(self, fname, device_encode=False):
        img = self.rgb_image_device() if device_encode else self.rgb_image()
        h, w, _ = img.shape
        raw = b"".join(b"\x00" + img[y].tobytes() for y in range(h))

        def chunk(tag, data):
            c = struct.pack(">I", len(data)) + tag + data
            return c + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)

        png = b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 2, 0, 0, 0)) + chunk(b"IDAT", zlib.compress(raw, 6)) + chunk(b"IEND", b"")
        with open(fname, "wb") as f:
            f.write(png)
