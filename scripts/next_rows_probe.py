import sys, json
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
print(json.dumps(bench.next_rows_leg(torch.device("cuda:0"), "bf16")))
