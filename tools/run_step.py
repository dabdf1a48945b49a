"""eager (no CUDA graph) FBCRNN train steps at the bench workload, for `ncu` launch lists."""
import sys
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
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    train.train_step(model, opt, batch)
torch.cuda.synchronize()
