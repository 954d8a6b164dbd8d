import os, sys; sys.path.insert(0, '.')
import numpy as np, torch
from jyutvoice_b200 import CausalConditionalCFM, CausalConditionalDecoder, synthetic
from oracle.make_golden import est_inputs
g = np.load(os.path.join('tests', 'golden', 'estimator_fwd.npz'))
cfm = CausalConditionalCFM(estimator=CausalConditionalDecoder(precision='bf16'))
cfm.load_state_dict(synthetic.make_estimator_state_dict(), strict=True)
cfm = cfm.cuda()
a = [z.cuda() for z in est_inputs(int(g['seed']), int(g['R']), int(g['T']), list(g['lens']))]
v = cfm.estimator(*a).cpu()
ref = torch.from_numpy(g['out'])
print('ERR', float((v - ref).abs().max()), float((v - ref).pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()))
