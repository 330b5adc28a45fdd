"""Fused grad-clip + AdamW kernel against torch.optim.AdamW + clip_grad_norm_ (the reference's
optimiser, configs/_base_/schedules/cosine_2x.py:1-9)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_flat_trainer_matches_torch_adamw():
    from geomae_b200.train import FlatTrainer
    torch.manual_seed(0)

    class Tiny(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.lin = torch.nn.Linear(37, 53)
            self.norm = torch.nn.LayerNorm(53)
            self.out = torch.nn.Linear(53, 5)

    ref, mine = Tiny().cuda(), Tiny().cuda()
    mine.load_state_dict(ref.state_dict())
    decay = [p for k, p in ref.named_parameters() if "norm" not in k]
    nodecay = [p for k, p in ref.named_parameters() if "norm" in k]
    opt = torch.optim.AdamW([dict(params=decay, weight_decay=0.05), dict(params=nodecay, weight_decay=0.0)],
                            lr=1e-2, betas=(0.9, 0.999), eps=1e-8)
    tr = FlatTrainer(mine, lr=1e-2, max_grad_norm=0.7)
    for it in range(5):
        x = torch.randn(64, 37, device="cuda") * (3.0 if it % 2 else 0.1)
        for m in (ref, mine):
            for p in m.parameters():
                if p.grad is not None:
                    p.grad.zero_()
            m.out(m.norm(m.lin(x))).pow(2).sum().backward()
        norm = torch.nn.utils.clip_grad_norm_(ref.parameters(), 0.7)
        opt.step()
        tr.optimizer_step()
        assert abs(float(tr.stats[0]) - float(norm)) <= 1e-5 * float(norm)
        for (k, a), (_, b) in zip(ref.named_parameters(), mine.named_parameters()):
            torch.testing.assert_close(b, a, rtol=2e-5, atol=2e-6, msg=k)


def test_flat_trainer_checkpoint_resume_and_binding_check():
    """state_dict / load_state_dict reproduce the run exactly; a detached parameter is reported, not silently skipped."""
    from geomae_b200.train import FlatTrainer

    def make():
        torch.manual_seed(1)
        m = torch.nn.Sequential(torch.nn.Linear(16, 32), torch.nn.LayerNorm(32), torch.nn.Linear(32, 4)).cuda()
        return m, FlatTrainer(m, lr=1e-2, max_grad_norm=1.0)

    def step(m, tr, seed):
        g = torch.Generator(device="cuda").manual_seed(seed)
        tr.zero_grad()
        m(torch.randn(8, 16, device="cuda", generator=g)).pow(2).sum().backward()
        tr.optimizer_step()

    ma, ta = make()
    for i in range(3):
        step(ma, ta, i)
    ckpt_model, ckpt_opt = {k: v.clone() for k, v in ma.state_dict().items()}, ta.state_dict()
    for i in range(3, 6):
        step(ma, ta, i)
    mb, tb = make()
    mb.load_state_dict(ckpt_model)          # copies in place: parameters stay views of the flat buffer
    tb.load_state_dict(ckpt_opt)
    tb.check_bindings()
    for i in range(3, 6):
        step(mb, tb, i)
    for (k, a), (_, b) in zip(ma.named_parameters(), mb.named_parameters()):
        assert torch.equal(a, b), k
    mb[0].weight.grad = None
    with pytest.raises(RuntimeError, match="no longer lives"):
        tb.check_bindings()
