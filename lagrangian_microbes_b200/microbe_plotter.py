"""MicrobePlotter: the reference's frame renderer with the scatter plot rasterised on the B200.

Mirrors /root/reference/microbe_plotter.py: same constructor arguments (:33-40), ``plot_frames(start_time, end_time, dt)``
(:66-80) and ``plot_frame(i, frame_time)`` (:82-155), reading ``microbe_data.nc`` from ``output_dir`` (the reference's
quirk Q4, :85: not from ``input_dir``) and writing ``lagrangian_microbes_{i:05d}.png`` there (:149-152).

What a frame is here: the reference's one ``plt.scatter`` call over every microbe, coloured by species (:132-146), as a
raster -- ``lm_rasterize`` bins the microbes of column i into width x height pixels over the plotted extent
(reference: lon -180..-120, lat 0..60 in PlateCarree, :102 -- 180..240 in the 0-360 convention the particle files use),
``lm_compose_frame`` colours each pixel like the microbe matplotlib draws last (array order) and the PNG is written
with zlib.  The map furniture of the reference (cartopy land polygons, grid lines, legend, title, :100-148) needs
matplotlib + cartopy + Natural Earth data, none of which exist here: it is not drawn.  ``microbe_marker_size`` (points^2
at 300 dpi in the reference) becomes the side of a square marker in pixels: ``round(sqrt(size) / 3)`` clamped to >= 1 --
the raster is computed at 1/marker of the output resolution and enlarged.  ``N_procs`` is accepted and ignored: the
frames of a run are rendered one after the other on the device (the reference fans them out with joblib, :76-80).
"""
import ctypes
import logging
import os

import numpy as np

from . import io as lmio
from .analysis import color_rgb
from .interactions import PAPER_COLOR, ROCK_COLOR, SCISSORS_COLOR

logger = logging.getLogger(__name__)


class MicrobePlotter:
    def __init__(
            self,
            N_procs=1,
            dark_theme=False,
            microbe_marker_size=10,
            input_dir=".",
            output_dir=".",
            extent=(180.0, 240.0, 0.0, 60.0),
            width=1600,
            height=900,
            mode="last_drawn",
    ):
        self.N_procs = N_procs
        self.dark_theme = dark_theme
        self.microbe_marker_size = microbe_marker_size
        self.input_dir = input_dir
        self.output_dir = output_dir
        self.extent = tuple(float(v) for v in extent)          # lon_min, lon_max, lat_min, lat_max
        self.marker_px = max(1, int(round(np.sqrt(float(microbe_marker_size)) / 3.0)))
        self.width, self.height = int(width), int(height)
        assert mode in ("last_drawn", "plurality")
        self.mode = mode
        background = color_rgb("black") if dark_theme else color_rgb("white")
        self.palette = np.array([background, color_rgb(ROCK_COLOR), color_rgb(PAPER_COLOR), color_rgb(SCISSORS_COLOR)],
                                dtype=np.uint8)
        self._data = None

    # ---- the device part ---------------------------------------------------------------------------
    def render(self, lons, lats, species, device=None):
        """uint8 (height, width, 3) frame of one snapshot (NumPy arrays or CUDA tensors)."""
        import torch
        from . import _lib
        if not torch.cuda.is_available():
            raise RuntimeError("MicrobePlotter needs a CUDA device: the rasteriser has no CPU fallback")
        L = _lib.lib()
        device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        w, h = max(1, self.width // self.marker_px), max(1, self.height // self.marker_px)
        with torch.cuda.device(device):
            lo = self._dev(lons, torch.float32, device)
            la = self._dev(lats, torch.float32, device)
            sp = self._dev(species, torch.int8, device)
            n = lo.numel()
            assert la.numel() == n and sp.numel() == n
            counts = torch.empty((3, h, w), dtype=torch.int32, device=device)
            top = torch.empty((h, w), dtype=torch.int32, device=device)
            rgb = torch.empty((h, w, 3), dtype=torch.uint8, device=device)
            stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
            p = ctypes.c_void_p
            _lib.check(L.lm_rasterize(p(lo.data_ptr()), p(la.data_ptr()), p(sp.data_ptr()), n, self.extent[0],
                                      self.extent[1], self.extent[2], self.extent[3], w, h, p(counts.data_ptr()),
                                      p(top.data_ptr()), stream), "lm_rasterize")
            mode = _lib.LM_FRAME_LAST_DRAWN if self.mode == "last_drawn" else _lib.LM_FRAME_PLURALITY
            _lib.check(L.lm_compose_frame(p(counts.data_ptr()), p(top.data_ptr()), p(sp.data_ptr()), w, h, mode,
                                          self.palette.tobytes(), p(rgb.data_ptr()), stream), "lm_compose_frame")
            img = rgb.cpu().numpy()
        if self.marker_px > 1:
            img = np.repeat(np.repeat(img, self.marker_px, axis=0), self.marker_px, axis=1)
        return img

    @staticmethod
    def _dev(x, dtype, device):
        import torch
        if isinstance(x, torch.Tensor):
            return x.to(device=device, dtype=dtype).contiguous()
        return torch.from_numpy(np.ascontiguousarray(x, dtype={torch.float32: np.float32, torch.int8: np.int8}[dtype])).to(device)

    # ---- the reference's interface -------------------------------------------------------------------
    def plot_frames(self, start_time, end_time, dt):
        iters = (end_time - start_time) // dt
        times = [start_time + n * dt for n in range(iters)]
        logger.info("Plotting {:d} frames from {:}->{:} on the GPU.".format(iters, start_time, end_time))
        for i, t in enumerate(times):
            self.plot_frame(i, t)

    def plot_frame(self, i, frame_time):
        if self._data is None:
            nc_input_filepath = os.path.join(self.output_dir, "microbe_data.nc")       # microbe_plotter.py:85 (Q4)
            self._data = lmio.read_particle_file(nc_input_filepath)
        microbe_data = self._data
        logger.info("Plotting frame {:d}...".format(i))
        microbe_lons = np.asarray(microbe_data["longitude"][:, i])
        microbe_lats = np.asarray(microbe_data["latitude"][:, i])
        species = np.asarray(microbe_data["species"][:, i])
        img = self.render(microbe_lons, microbe_lats, species)
        png_filename = "lagrangian_microbes_" + str(i).zfill(5) + ".png"
        png_filepath = os.path.join(self.output_dir, png_filename)
        logger.info("Saving figure: {:s}".format(png_filepath))
        lmio.write_png(png_filepath, img)
        return png_filepath
