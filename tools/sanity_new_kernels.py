import numpy as np, torch, sys, os
sys.path.insert(0, os.getcwd())
from qublas_b200 import ops
from qublas_b200.qtypes import QType, RND_CONV, RND_POS_INF, SAT_TCPL, TREE_VARIADIC
from oracle import oracle
from tests.helpers import reduce_plan
# C3 tile kernel (substituted operands, negated S form), a few instances incl. minima
a = QType(7, 8); mul = QType(14, 8, True, RND_CONV, SAT_TCPL)
lv = [QType(15 + l, 8 - l, True, RND_POS_INF, SAT_TCPL) for l in range(6)]
c = QType(17, 4, True, RND_POS_INF, SAT_TCPL)
plan = ops.make_gemul_plan(a, a, c, 64, 64, 64, mul=mul, add=lv)
b = 5
A = torch.empty(b * 4096, dtype=torch.int32, device="cuda"); B = torch.empty_like(A); Cd = torch.empty_like(A)
ops.synth_fill(A, a.raw_min, a.raw_max, 1, 0); ops.synth_fill(B, a.raw_min, a.raw_max, 1, 1)
A[:100] = a.raw_min; B[:100] = a.raw_min
ops.Qgemul(Cd, A, B, plan, batch=b)
want = oracle.gemul_batched(plan, A.cpu().numpy(), B.cpu().numpy(), b)
assert np.array_equal(Cd.cpu().numpy(), want); print("c3 ok", ops.plan_info_gemul(plan).kernel)
print("err", ops.gemul_error(Cd, A, B, plan, batch=b)["sqnr_db"])
# fused narrowing
M, N, K = 256, 256, 1152
t = QType(7, 0); acc = QType(30, 0)
p5 = ops.make_gemul_plan(t, t, acc, M, N, K, mul=QType(15, 0), add=[acc], trans_a=True)
rng = np.random.default_rng(0)
A5 = torch.from_numpy(rng.integers(-128, 128, size=M * K).astype(np.int32)).cuda(); B5 = torch.from_numpy(rng.integers(-128, 128, size=K * N).astype(np.int32)).cuda()
C5 = torch.empty(M * N, dtype=torch.int32, device="cuda")
ops.Qgemul(C5, A5, B5, p5)
ref = (A5.view(M, K).double() @ B5.view(N, K).t().double()).t().contiguous().view(-1)
assert torch.equal(C5.double(), ref); print("c5 fused ok")
# variadic reduce
elem = QType(8, 8); rl = [QType(9 + l, 8 - l, True, RND_POS_INF, SAT_TCPL) for l in range(8)] + [QType(17 + (l - 6) // 2, 0, True, RND_POS_INF, SAT_TCPL) for l in range(8, 20)]
n = 70001
x = rng.integers(elem.raw_min, elem.raw_max + 1, size=3 * n).astype(np.int32)
out, _ = ops.Qreduce(torch.from_numpy(x).cuda(), elem, rl, length=n, batch=3, form=TREE_VARIADIC)
w, _ = oracle.reduce_batched(reduce_plan(elem, rl, n, form=TREE_VARIADIC), x, 3)
assert np.array_equal(out.cpu().numpy(), w); print("var reduce ok")
