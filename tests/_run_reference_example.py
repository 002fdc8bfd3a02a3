"""Run an UNMODIFIED example script of the reference (its examples are its tests) against this
package on a CPU host:

    torchrun --nproc-per-node 2 tests/_run_reference_example.py /path/to/reference/examples/x.py

* ``compat.install_alias()`` makes the script's ``from torchdistpackage... import`` lines resolve
  to this package;
* ``.cuda()`` is the identity and ``torch.cuda.synchronize`` a no-op (the scripts move everything
  to the GPU; here the same code runs on gloo);
* modules the scripts import but do not need for the path under test (``timm``) are stubbed.
Nothing of the script itself is edited."""
import os
import runpy
import sys
import types

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import torchdistpackage_b200.compat as compat  # noqa: E402

assert compat.install_alias()
if not torch.cuda.is_available():
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self
    torch.cuda.synchronize = lambda *a, **k: None
    torch.cuda.set_device = lambda *a, **k: None
if "timm" not in sys.modules:
    try:
        import timm  # noqa: F401
    except Exception:
        # timm is not in this image: its resnet50 is torchvision's architecture
        stub = types.ModuleType("timm")

        def create_model(name, pretrained=False, **kw):
            import torchvision.models as tvm
            if pretrained or not hasattr(tvm, name):
                raise RuntimeError(f"timm is not installed (asked for {name})")
            return getattr(tvm, name)()
        stub.create_model = create_model
        sys.modules["timm"] = stub
import pdb  # noqa: E402

pdb.set_trace = lambda *a, **k: None      # the scripts drop into pdb before failing an assert

script = sys.argv[1]
sys.argv = [script] + sys.argv[2:]
runpy.run_path(script, run_name="__main__")
import torch.distributed as dist  # noqa: E402

if dist.is_initialized():
    dist.barrier()
    dist.destroy_process_group()
print("REFERENCE_EXAMPLE_DONE", flush=True)
