"""The DCUE trainer API (reference dcrecommend/nn/dcue.py) on the B200 kernels, driven with tiny
synthetic datasets: train/eval epochs against the oracle, factor extraction, predict/score,
top-k recommendation, save/load."""
import importlib
import os

import numpy as np
import pytest
import torch
from torch.utils.data import DataLoader

from oracle import dcue_oracle as O
from tests.synthetic_data import SynthItemSet, SynthPredSet, SynthTrainSet, SynthWorld

pkg = importlib.import_module("amplifai-deepcontentrecommenders_b200")
pytestmark = pytest.mark.gpu


def _trainer(world, **kw):
    t = pkg.DCUE(batch_size=8, neg_batch_size=world.negs, lr=1e-4, num_epochs=1, eval_pct=1.0, **kw)
    t.num_workers = 0
    t.n_users, t.n_items = world.n_users, world.n_songs
    t.epoch_size = 16
    return t


def test_train_and_eval_epoch_match_oracle():
    w = SynthWorld()
    ds = SynthTrainSet(w, w.pairs[:16])
    loader = DataLoader(ds, batch_size=8, shuffle=False, drop_last=True)
    torch.manual_seed(0)
    t = _trainer(w)
    t._init_nn()
    t.scheduler.step()
    params = {k: v.detach().cpu().clone() for k, v in t.model.state_dict().items()}
    n, loss = t._train_epoch(loader)
    assert n == 16
    # oracle: same two Adam steps with the scheduler's learning rates
    names = [k for k, v in params.items() if v.is_floating_point() and "running_" not in k]
    leaves = [torch.nn.Parameter(params[k].clone()) for k in names]
    opt = torch.optim.Adam(leaves, 1e-4, (0.9, 0.99), 1e-8, 0)
    sched = importlib.import_module("amplifai-deepcontentrecommenders_b200.optim").CyclicLRWithRestarts(
        opt, 8, epoch_size=16, restart_period=30, t_mult=2, policy='cosine')
    sched.step()
    tot = 0.0
    for batch in loader:
        cur = dict(params)
        cur.update({k: l.detach() for k, l in zip(names, leaves)})
        r = O.train_step_grads(cur, batch['u'], batch['X'], batch['Ns'], t.model_type, t.margin,
                               operand_dtype=torch.float16, grad_dtype="fp16_scaled")
        for l, k in zip(leaves, names):
            l.grad = r["grads"][k].float()
        opt.step()
        sched.batch_step()
        params.update(r["new_stats"])
        tot += r["loss"].item() * 8
    assert abs(loss - tot / 16) < 2e-3 * abs(tot / 16)
    n2, vloss = t._eval_epoch(loader)
    assert n2 == 16 and np.isfinite(vloss)
    assert t.optimizer.param_groups[0]['lr'] == opt.param_groups[0]['lr']


def test_factors_predict_score_topk_and_checkpoint(tmp_path):
    w = SynthWorld()
    torch.manual_seed(1)
    t = _trainer(w)
    t._init_nn()
    items = SynthItemSet(w)
    t._user_factors(items)
    t._item_factors(items, n_iter=2)
    assert t.user_factors.shape == (w.n_users, 100) and t.item_factors.shape == (w.n_songs, 100)
    # factors == eval-mode oracle towers
    p = {k: v.detach().cpu() for k, v in t.model.state_dict().items()}
    with torch.no_grad():
        uf = O.user_forward(p, torch.arange(w.n_users))
        itf = O.tower_forward(p, w.mels, t.model_type, training=False, operand_dtype=torch.float16)
    assert (t.user_factors.cpu() - uf).abs().max() < 1e-5 * uf.abs().max()
    assert (t.item_factors.cpu() - itf).abs().max() < 3e-3 * itf.abs().max()
    # predict: cosine of factor rows for every candidate of one user
    pred = SynthPredSet(w)
    loader = DataLoader(pred, batch_size=16, shuffle=False)
    scores, targets = t.predict(3, loader)
    cand = [s for _, s, _ in pred.rows]
    ref = O.cosine(t.user_factors.cpu()[3].expand(len(cand), -1), t.item_factors.cpu()[cand])
    assert np.allclose(scores, ref.numpy(), atol=1e-5) and targets == [y for _, _, y in pred.rows]
    # score / score_song (device AUC + AP kernel) against the reference's estimator evaluated with sklearn on the
    # predict() lists (dcrecommend/nn/dcue.py:380-476), pred and truth being different splits
    from sklearn.metrics import average_precision_score, roc_auc_score
    truth_loader = DataLoader(SynthPredSet(w, w.pairs[::2]), batch_size=16)
    pred2 = SynthPredSet(w, w.pairs[1::2])
    loader2 = DataLoader(pred2, batch_size=16, shuffle=False)
    users = [0, 1, 2, 5]
    auc, mAP = t.score(users, loader2, truth_loader)
    ref_auc, ref_map = [], []
    for user in users:
        sp, tp = (np.array(x) for x in t.predict(user, loader2))
        st, tt = (np.array(x) for x in t.predict(user, truth_loader))
        halves = [(list(sp[tp == 1]) + list(st[tt == 0]), list(tp[tp == 1]) + list(tt[tt == 0])),
                  (list(sp[tp == 0]) + list(st[tt == 1]), list(tp[tp == 0]) + list(tt[tt == 1]))]
        total = len(halves[0][0]) + len(halves[1][0])
        part = [1 if sum(tg) == len(tg) else 0 if sum(tg) == 0 else roc_auc_score(tg, sc) for sc, tg in halves]
        ref_auc.append(len(halves[0][0]) / total * part[0] + len(halves[1][0]) / total * part[1])
        ref_map.append(average_precision_score(halves[0][1] + halves[1][1], halves[0][0] + halves[1][0]))
    assert abs(auc - np.mean(ref_auc)) < 1e-9 and abs(mAP - np.mean(ref_map)) < 1e-9
    songs = pred2.uniq_songs[:4]
    sauc, smap = t.score_song(songs, loader2)
    ra, rm = [], []
    for song in songs:
        sc, tg = t.predict_song(song, loader2)
        ra.append(1 if sum(tg) == len(tg) else 0 if sum(tg) == 0 else roc_auc_score(tg, sc))
        rm.append(1 if sum(tg) == len(tg) else 0 if sum(tg) == 0 else average_precision_score(tg, sc))
    assert abs(sauc - np.mean(ra)) < 1e-9 and abs(smap - np.mean(rm)) < 1e-9
    # top-k recommendation vs the oracle's all-pairs scorer
    ts, ti = t.recommend_topk(k=5)
    vs, vi = O.topk_scores(t.user_factors.cpu(), t.item_factors.cpu(), 5)
    assert (ts.cpu() - vs).abs().max() < 2e-3
    agree = np.mean([len(set(a.tolist()) & set(b.tolist())) / 5 for a, b in zip(ti.cpu(), vi)])
    assert agree > 0.95
    # checkpoint round trip
    t.nn_epoch = 4
    t.save(models_dir=str(tmp_path))
    t2 = pkg.DCUE()
    t2.load(os.path.join(str(tmp_path), t._format_model_subdir()), 4)
    assert t2.nn_epoch == 5 and t2.batch_size == 8
    for k, v in t.model.state_dict().items():
        assert torch.equal(v.cpu(), t2.model.state_dict()[k].cpu()), k
    assert torch.equal(t2.user_factors.cpu(), t.user_factors.cpu())


def test_fit_runs_end_to_end(tmp_path):
    w = SynthWorld(n_users=20, n_songs=20)   # 120 (user, song) pairs -> epoch_size 8 (one batch per sub-epoch)
    tr, va = SynthTrainSet(w, w.pairs[:96]), SynthTrainSet(w, w.pairs[96:])
    t = pkg.DCUE(batch_size=8, neg_batch_size=w.negs, lr=1e-4, num_epochs=1, eval_pct=1.0)
    t.num_workers = 0
    np.random.seed(0)
    t.fit(tr, va, va, SynthPredSet(w, w.pairs[96:]), SynthPredSet(w, w.pairs[:96]), SynthItemSet(w), w.n_users,
          w.n_songs, "triplets.txt", "metadata.csv", str(tmp_path))
    # like the reference, a started sweep over the 10 sub-epoch loaders always finishes (nn_epoch 0 = evaluation only)
    assert t.nn_epoch == 10 and t.user_factors is not None
    assert os.listdir(os.path.join(str(tmp_path), t._format_model_subdir()))
