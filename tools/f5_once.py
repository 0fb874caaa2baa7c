"""Two config-3 utterances (U from argv, default 1) through the engine, graphs off: the command the ncu launch list / full captures
of round 2 profile (tools/ncu_r02.sh). Prints the launch count of the second utterance."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import b200tts  # noqa: E402,F401
from b200tts import capi, config, synth, weights  # noqa: E402


def main():
    import torch
    U = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else -1
    eng = capi.Engine(0)
    eng.set_option("cuda_graphs", 0)
    cfg = config.F5
    dsd = synth.f5_dit_state(4321)
    eng.load_state("dit", weights.dit_engine_tensors(dsd, cfg))
    eng.load_state("vocos", weights.vocos_engine_tensors(synth.vocos_state(2468), cfg))
    eng.load_state("f5", weights.f5_export_constants(dsd, cfg))
    eng.f5_build()
    L, n_text = 144000, 150
    ins = [synth.f5_inputs(1000 + i, L, n_text) for i in range(U)]
    N = int(ins[0][2][0])
    ns = 256 * (N - (L // 256 + 1) - 1)
    audio = torch.from_numpy(np.stack([a.reshape(-1) for a, _, _, _ in ins])).cuda()
    ids = torch.from_numpy(np.stack([t.reshape(-1) for _, t, _, _ in ins])).cuda()
    noise = torch.from_numpy(np.stack([n.reshape(-1) for _, _, _, n in ins])).cuda()
    pcm = torch.zeros((U, ns), dtype=torch.int16, device="cuda")
    for i in range(2):
        l0 = eng.launch_count()
        eng.f5_synthesize_batch_device(U, audio.data_ptr(), L, ids.data_ptr(), n_text, N, noise.data_ptr(), pcm.data_ptr(), precision=capi.F16,
                                       n_steps=steps)
        eng.synchronize()
        print(f"utterance batch {i}: {eng.launch_count() - l0} launches", flush=True)


if __name__ == "__main__":
    main()
