import numpy as np, torch, sys
sys.path.insert(0,'.')
from env_build_b200 import _lib, synthetic as syn
from env_build_b200.dynamics_and_models import EnvironmentModel
from env_build_b200.endtoend_env_utils import VEHICLE_MODE_LIST
rng=np.random.default_rng(0)
B,V=1024,32
m=EnvironmentModel('left',0,mode='training',veh_mode_list=syn.tiled_mode_list(VEHICLE_MODE_LIST['left'],V))
ref=syn.make_ref_indexes(rng,B); obs=syn.make_obs(rng,B,'left',V,m.ref_path.path_list,ref)
m.reset(obs,ref)
r=m.rollout_out(syn.make_actions(rng,1,B)[0])
torch.cuda.synchronize()
print('ok', float(r[1].sum()))
