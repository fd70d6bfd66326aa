"""Imports the reference's own fixture tensors (assets/*.safetensors, produced by python-reference/scripts/extract_refs.py
and extract_decoder_refs.py) into tests/golden/ref_assets.npz so that the GPU box, which has no /root/reference, can use
them:
  voice_conditioning [87,1024]   the real conditioning rows of assets/ref.wav (parity_tests.rs:60-142 pins them at 2e-2)
  mimi_input [167040]            the 24 kHz PCM the reference feeds its Mimi encoder for that voice (= 87 frames)
  decoder stages                 latent_from_flowlm / denormalized / quantized / after_upsample / after_decoder_transformer /
                                 final_audio of one frame with the real checkpoint (parity_tests.rs:521-612: 0.05/0.05/0.1)
Run here (the container with /root/reference):  python tests/golden/import_reference_fixtures.py"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from pocket_tts_b200.tts_model import read_safetensors  # noqa: E402

A = Path("/root/reference/assets")
vc = read_safetensors(A / "ref_voice_conditioning.safetensors")["voice_conditioning"]
mi = read_safetensors(A / "ref_mimi_input.safetensors")["mimi_input"]
dec = read_safetensors(A / "ref_decoder_intermediates.safetensors")
out = {"voice_conditioning": vc.reshape(87, 1024).astype(np.float32), "mimi_input": mi.reshape(-1).astype(np.float32)}
out.update({f"dec_{k}": v.astype(np.float32) for k, v in dec.items()})
np.savez_compressed(ROOT / "tests" / "golden" / "ref_assets.npz", **out)
print({k: v.shape for k, v in out.items()})
