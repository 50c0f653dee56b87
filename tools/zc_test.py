import sys, os
sys.path.insert(0, '.')
import torch, bsuite_b200
for track in (False, True):
    for B in (4096, 65536):
        env = bsuite_b200.load_from_id('deep_sea/11', batch=B, device='cuda', seed=0, track_episodes=track)
        ring = [env.make_buffers() for _ in range(2)]
        acts = torch.randint(0, 2, (8, B), dtype=torch.int32).pin_memory()
        hb = env.make_host_buffers()
        for i in range(6):
            env.step_host(acts[i % 8], hb, out=ring[i % 2])
        torch.cuda.synchronize()
        print('ok', track, B, flush=True)
        env.close()
