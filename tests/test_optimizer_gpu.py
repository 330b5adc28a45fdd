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
