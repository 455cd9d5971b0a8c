"""Timeline of one layer-chain CTA (SRHEP_CHAIN_DBG=1): one forward on synthetic events."""
import os, sys
os.environ["SRHEP_CHAIN_DBG"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from superresolutionhep_b200 import FlowModel
from superresolutionhep_b200.default_configs import flow_config
from superresolutionhep_b200.synthetic import synthetic_events, synthetic_noise, synthetic_state_dict
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
m = FlowModel(flow_config("single_e"), precision=os.environ.get("SRHEP_DBG_PRECISION", "fp16")); m.load_state_dict(synthetic_state_dict(m.dims, seed=7)); m.eval().cuda()
b = synthetic_events("single_e", B, seed=1234); x = synthetic_noise(b, seed=0)
db = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in b.items()}
for _ in range(2):
    m(db, x.cuda(), torch.full((B,), 0.5).cuda())
torch.cuda.synchronize()
