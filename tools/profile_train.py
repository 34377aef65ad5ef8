"""One eager bs16 train step (forward + loss + backward + clip + SGD) between cudaProfilerStart/Stop, after two warm-up steps:
  ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/train_step_launches.csv \
      python tools/profile_train.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    from transception_b200 import MSTransception
    from transception_b200.losses import CeDiceLoss
    dev = torch.device("cuda", 0)
    torch.manual_seed(1234)
    net = MSTransception(num_classes=9).to(dev).train()
    crit = CeDiceLoss(9)
    from transception_b200.optim import FusedSGD
    opt = FusedSGD(net.parameters(), lr=0.05, momentum=0.9, weight_decay=1e-4)
    g = torch.Generator().manual_seed(0)
    x = (torch.rand(16, 1, 224, 224, generator=g) * 2 - 1).to(dev)
    labels = torch.randint(0, 9, (16, 224, 224), generator=g).to(dev)

    def step():
        opt.zero_grad(set_to_none=True)
        loss = crit(net(x), labels)
        loss.backward()
        opt.step()
        return loss

    for _ in range(2):
        step()
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    loss = step()
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
    print("loss %.6f" % loss.item())


if __name__ == "__main__":
    main()
