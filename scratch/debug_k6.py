import sys, numpy as np, torch
sys.path.insert(0, '.')
from tests import goldutil as G, gpu_harness as H
from tests.test_gpu_tracker_train import _replay
from oracle import nets
name = sys.argv[1] if len(sys.argv) > 1 else "kuaishou_N5"
z = G.load(name); c = G.cfg(z)
trk = H.make_tracker(z, c); pol = H.make_policy(z, c, None)
buf, _ = _replay(H, z, c, 0, trk, pol)
buf.sync_device()
B, L, S = c["B"], buf.sub_size, 20
lens = buf._lengths
print("lens", lens, "L", L, "users", buf.d_users.cpu().numpy(), z["it0/users"])
rng = np.random.default_rng(0)
d_obs = np.zeros((B * L, S), dtype=np.float32)
for e in range(B):
    d_obs[e * L:e * L + lens[e]] = rng.normal(size=(lens[e], S))
check = torch.zeros(B * L, S, device="cuda")
trk.zero_grad()
trk.backward_from_buffer(buf, torch.tensor(d_obs, device="cuda"), buf.d_users, obs_check=check)
torch.cuda.synchronize()
idx = buf.sample_index(0); it = torch.as_tensor(idx, device="cuda")
a, b = check[it].cpu().numpy(), buf.obs[it].cpu().numpy()
print("forward max abs err", np.abs(a - b).max(), "scale", np.abs(b).max())
P = {k: v.clone().requires_grad_(k != "pos_encoder.pe") for k, v in nets.to_params(z, "init/tracker/").items()}
users, acts, rews = z["it0/users"], buf.act.reshape(B, L), buf.rew.reshape(B, L)
loss = 0.0
for e in range(B):
    n = int(lens[e])
    toks = [nets.user_token(P, users=[users[e]])]
    if n > 1:
        toks.append(nets.action_token(P, rews[e, :n - 1], acts=acts[e, :n - 1]))
    X = torch.cat(toks, 0).unsqueeze(1)
    s = nets.encode(X, P, c["nhead"], all_positions=True)[:, 0]
    loss = loss + (s * torch.tensor(d_obs[e * L:e * L + n])).sum()
loss.backward()
mine = trk.layout.unpack(trk.grad)
for k, p in P.items():
    if k == "pos_encoder.pe": continue
    ref = p.grad.numpy(); m = mine[k].numpy()
    print(f"{k:60s} ref_max {np.abs(ref).max():.3e} mine_max {np.abs(m).max():.3e} max_err {np.abs(m-ref).max():.3e}")
