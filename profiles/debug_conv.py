import sys, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import test_gpu_conv_tc as T
for (N, Cin, Cout, sp) in [(1,16,16,(4,16,8)), (1,32,96,(6,12,12)), (1,32,112,(6,12,12)), (1,32,128,(6,12,12)), (1,32,128,(8,16,8)), (1,32,128,(1,16,8)), (1,64,80,(8,16,8))]:
    try:
        T._run(N, Cin, Cout, sp, bias=False)
        torch.cuda.synchronize()
        print('ok', N, Cin, Cout, sp, flush=True)
    except Exception as e:
        print('FAIL', N, Cin, Cout, sp, str(e)[:300], flush=True)
        break
