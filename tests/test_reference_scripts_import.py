"""The reference's own driver scripts import the drop-in unmodified when the drop-in directory precedes the
reference root on sys.path.  Runs only where /root/reference exists (build container)."""
import os
import subprocess
import sys

import pytest

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DROPIN = os.path.join(ROOT, "fast_neural_style_transfer_b200", "dropin")


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present on this machine")
def test_train_and_inference_scripts_resolve_to_dropin():
    code = (
        "import sys, inspect\n"
        "import train, inference\n"
        "import models.model, models.vgg19_net, losses.losses\n"
        "for m in (models.model, models.vgg19_net, losses.losses):\n"
        "    assert 'fast_neural_style_transfer_b200' in inspect.getfile(m), inspect.getfile(m)\n"
        "assert train.StyleTransferNet is models.model.StyleTransferNet\n"
        "assert train.gram_matrix is losses.losses.gram_matrix\n"
        "assert inference.StyleTransferNet is models.model.StyleTransferNet\n"
        "assert '/root/reference' in inspect.getfile(train.Dataset)\n"
        "net = train.StyleTransferNet(); assert len(net.state_dict()) == 58\n"
        "print('ok')\n")
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([DROPIN, REF]))
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, cwd=REF, timeout=300)
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr[-2000:]
