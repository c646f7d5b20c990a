import sys, numpy as np
sys.path.insert(0,'oracle'); sys.path.insert(0,'.')
import oracle, lpvmpc_b200 as lp
W=lp.workloads; m=lp.Map("L_shape")
rng=np.random.default_rng(21); B,T,N=40,12,40
x0=np.stack([rng.uniform(1.0,2.0,B),rng.normal(0,0.01,B),rng.normal(0,0.05,B),rng.normal(0,0.02,B),rng.normal(0,0.02,B)],axis=1)
s0=rng.uniform(0,18,B); x0[0]=[1,0,0,0,0]; s0[0]=0; x0[1,0]=0.5
cfg=oracle.make_cfg("planner",N,W.PLAN_DT,W.PLAN["Q"],W.PLAN["R"],W.PLAN["dR"],m.PointAndTangent,L_cf=W.PLAN["L_cf"])
fleet=lp.PlannerFleet(m,N=N,max_fleet=64,max_ey=0.2); fleet.start(x0,s0)
ref=oracle.plan_loop_state(x0,N,s0)
for t in range(T):
    fleet.run(1); got=fleet.read()
    oracle.plan_loop_run(cfg,oracle.default_settings(polish=1),ref,1,max_ey=0.2,threads=8)
    al=got["ctr"][:,3]==0
    print(t, "ctr eq", np.array_equal(got["ctr"],ref["ctr"]), "dSS %.2e dx %.2e du %.2e"%(np.abs(got["SS"][al]-ref["SS"][al]).max(), np.abs(got["x_pred"][al]-ref["x_pred"][al]).max(), np.abs(got["u_pred"][al]-ref["u_pred"][al]).max()), "worst b", np.abs(got["x_pred"][al]-ref["x_pred"][al]).reshape(al.sum(),-1).max(1).argmax())
