import sys, numpy as np, torch
sys.path.insert(0, '.')
from beer_b200 import ops, synthetic
from beer_b200.engine import EmissionParams, Utterances, VBEngine, WeightGroup
from beer_b200.synthetic import CONFIGS
dev = torch.device('cuda', 0)
c = dict(CONFIGS['cfg3']); U = 96; T = c['n_frames']; P, S, C, D = c['n_units'], c['n_states'], c['n_comp'], c['dim']
K, M = P * S, P * S * C
graph, _, _ = synthetic.phone_loop_graph(P, S)
plan = ops.GraphPlan(graph.init_log_probs.numpy(), graph.final_log_probs.numpy(), graph.trans_log_probs.numpy(), graph.pdf_id_mapping, n_pdfs=K)
means = 2.0 * torch.randn(K, D, generator=torch.Generator().manual_seed(0))
X = synthetic.sample_utterances(graph, means, U, T, seed=1, device=dev)
prior, post = synthetic.initial_normal_gamma(M, D, seed=2, device=dev)
conc = torch.full((K, C), 1.0 / C, device=dev)
em = EmissionParams(prior, post, comp_off=np.arange(K + 1) * C, weight_groups=(WeightGroup(0, K, C, conc.clone(), conc.clone()),))
eng = VBEngine(em, plan, Utterances(X, [T] * U), datasize=float(U * T), distributed=False)
off = torch.arange(U + 1, device=dev) * T
for it in range(14):
    # activity of (64-frame tile x 16-pdf block) pairs under the CURRENT model, as the statistics kernel would see it
    W, bias, ref = em.refresh(pack_tc=False)
    images = eng._images[0]
    eng.mix16.pack(W, bias, images['alpha'])
    llh2 = eng.mix16.emission(images)
    fref = eng.mix16.frame_ref(X, ref)
    lp = torch.empty(U * T, K, device=dev)
    ops.hmm_forward_backward(plan, llh2, fref, off, want_pdf_post=False, out_pdf_lpost=lp, llh_log2=True)
    lpp = torch.nn.functional.pad(lp, (0, (-K) % 16), value=float('-inf'))
    blk = lpp.reshape(U * T // 64, 64, lpp.shape[1] // 16, 16).amax(dim=(1, 3))
    frac = [(blk >= thr).float().mean().item() for thr in (-41.0, -60.0, -100.0)]
    elbo = float(eng.step().item())
    print(it, 'active fraction at thr -41/-60/-100:', [round(f, 4) for f in frac], 'elbo/frame', elbo / (U * U * T))
