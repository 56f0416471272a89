"""The torch-side host layer (Engine, FusedSimulation.run_to_file plain and packed, record.DeltaRecordPacker) driven
against the CPU emulator of the kernels (tests/cuda_emu) in a subprocess: tools/host_layer_on_emulator.py monkeypatches
torch.cuda inside its own process and needs ``python -O`` (the wrappers' is_cuda asserts).  Test infrastructure only:
the product path itself has no CPU mode."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_host_layer_against_the_emulated_kernels():
    r = subprocess.run([sys.executable, "-O", os.path.join(ROOT, "tools", "host_layer_on_emulator.py")],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    for line in ("run_to_file ok:", "delta record stream ok:", "run_to_file(packed=True) ok:", "packed stream decoded on read ok"):
        assert line in r.stdout, r.stdout[-2000:]
