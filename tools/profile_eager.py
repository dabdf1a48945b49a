"""host-side cost of the un-graphed (drop-in) train step: cProfile over eager steps at the bench workload.
    python tools/profile_eager.py [steps] > gpurun_out/eager_profile.txt"""
import cProfile
import io
import pstats
import sys
import time
sys.path.insert(0, '.')
import torch
import bench
from pb_sed_b200 import config, train
from pb_sed_b200.models import weak_label
B = 32
torch.manual_seed(0)
model = weak_label.CRNN.from_config_dict(config.fbcrnn_config(num_events=bench.NUM_EVENTS)).cuda()
model.emit_buffers = False
opt = train.Adam(model, lr=5e-4)
audio, weak, boundary = bench.synthetic_clips(B, 1234)
batch = {'audio_data': torch.from_numpy(audio).cuda(), 'weak_targets': torch.from_numpy(weak).cuda(),
         'boundary_targets': torch.from_numpy(boundary).cuda(), 'seq_len': [bench.T_FRAMES] * B}
n = int(sys.argv[1]) if len(sys.argv) > 1 else 5
for _ in range(3):
    train.train_step(model, opt, batch)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(n):
    train.train_step(model, opt, batch)
t_issue = time.perf_counter() - t0
torch.cuda.synchronize()
t_all = time.perf_counter() - t0
print(f'eager step: host issue time {1e3 * t_issue / n:.2f} ms, wall (incl. device drain) {1e3 * t_all / n:.2f} ms')
pr = cProfile.Profile()
pr.enable()
for _ in range(n):
    train.train_step(model, opt, batch)
pr.disable()
torch.cuda.synchronize()
buf = io.StringIO()
pstats.Stats(pr, stream=buf).sort_stats('cumulative').print_stats(45)
print(buf.getvalue())
buf = io.StringIO()
pstats.Stats(pr, stream=buf).sort_stats('tottime').print_stats(30)
print(buf.getvalue())
