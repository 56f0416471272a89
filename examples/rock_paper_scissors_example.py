"""The reference's end-to-end example (/root/reference/rock_paper_scissors_example.py:14-36) on the B200 path: the
same calls with the same arguments -- only the imports change.  With --plots the frames of
:38-42 are rendered too (MicrobePlotter: scatter rasterised on the GPU, PNGs without the cartopy map furniture), plus
the species census figure of analysis.py and the per-species pair-distance histogram of the reference's sandbox; the
ffmpeg movie (:44-53) is out of scope.  Velocities are the synthetic OSCAR-like field (no network here); register a reader of
the real `oscar_vel<year>.nc` with `velocity_fields.register_dataset_provider` to use the product itself.

    python examples/rock_paper_scissors_example.py [output_dir] [--plots]
"""
import os
import sys
from datetime import datetime, timedelta

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from lagrangian_microbes_b200 import (InteractionSimulator, ParticleAdvecter, rock_paper_scissors,   # noqa: E402
                                      uniform_particle_locations)
from lagrangian_microbes_b200 import io as lmio                                                          # noqa: E402

N = 10000  # Number of particles
_args = [a for a in sys.argv[1:] if not a.startswith("--")]
output_dir = _args[0] if _args else "lagrangian_microbes_output"

start_time = datetime(2017, 1, 1)
end_time = datetime(2017, 2, 1)
dt = timedelta(hours=1)

# Generate initial locations for each particle.
particle_lons, particle_lats = uniform_particle_locations(N_particles=N, lat_min=25, lat_max=35, lon_min=205, lon_max=215)

# Create a particle advecter that will the advect the particles we just generated (4 output tiles).
pa = ParticleAdvecter(particle_lons, particle_lats, N_procs=4, velocity_field="OSCAR", output_dir=output_dir, Kh=100)

# Advect the particles and save all the data to NetCDF.
pa.time_step(start_time, end_time, dt)
pa.create_netcdf_file(start_time, end_time, dt)

# Create an interaction simulator that uses the rock-paper-scissors pair interaction.
rps_interaction = rock_paper_scissors(N_microbes=N, pRS=0.5, pPR=0.5, pSP=0.5)
# (advection_dir defaults to "." as in the reference, interaction_simulator.py:29: the reference script only works
# when run from inside output_dir; say where particle_data.nc is.)
isim = InteractionSimulator(pair_interaction=rps_interaction, interaction_radius=0.05, advection_dir=output_dir,
                            output_dir=output_dir)

# Simulate the interactions.
isim.time_step(start_time, end_time, dt)

m = lmio.read_particle_file(os.path.join(output_dir, "microbe_data.nc"))
sp = m["species"]
print("steps:", sp.shape[1], " species counts at the first / last step:",
      [int((sp[:, 0] == k).sum()) for k in (1, 2, 3)], [int((sp[:, -1] == k).sum()) for k in (1, 2, 3)])

if "--plots" in sys.argv:
    from lagrangian_microbes_b200 import analysis
    from lagrangian_microbes_b200.microbe_plotter import MicrobePlotter

    # Create a microbe plotter that will produce a plot of all the microbes at a single iteration (:38-42).
    mp = MicrobePlotter(N_procs=-1, dark_theme=True, input_dir=output_dir, output_dir=output_dir,
                        extent=(200.0, 220.0, 20.0, 40.0))
    mp.plot_frames(start_time, start_time + 100 * dt, dt)          # "Plot the first 100 frames and save them to disk."
    print("species census figure:", analysis.species_count_figure(output_dir, start_time, end_time, dt))
    last = sp.shape[1] - 1
    hists = analysis.species_pair_distance_histograms(m["longitude"][:, last], m["latitude"][:, last], sp[:, last], bins=70)
    for k, name in ((1, "rocks"), (2, "papers"), (3, "scissors")):
        peak = int(hists[k][:71].argmax())
        print("%-8s %9d pairs, most in the bin around %.0f m" % (name, int(hists[k].sum()), 10 ** (peak / 10)))
