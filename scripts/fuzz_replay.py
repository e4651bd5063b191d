"""Randomised check of the planner + tile-kernel logic + support tracking through the CPU replay (no GPU):
random dense and sparse circuits at 12-18 qubits on 1 / 2 / 4 / 8 emulated ranks against the oracle, with a random
store-side mode (swap rounds on loads / the restore on a store / every round on a store) and tail-deferral threshold.
Usage: python scripts/fuzz_replay.py [seed] [seconds]   (round 1: 414 circuits on 4 seeds x 150 s; round 2, with the store modes, thresholds, home-going and split swap rounds: 1073 circuits on 8 seeds x 900 s)"""
import sys, time
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from damavand_b200 import circuits
from oracle.oracle import OracleCircuit
from tests.helpers import emu_run, rel_err
rng=np.random.default_rng(int(sys.argv[1]) if len(sys.argv)>1 else 0)
t0=time.time(); n_ok=0
while time.time()-t0 < float(sys.argv[2] if len(sys.argv)>2 else 120):
    n=int(rng.integers(12,19)); world=int(rng.choice([1,2,4,8]));
    while n-(world.bit_length()-1) < 12: world//=2
    gates=int(rng.integers(1,120 if rng.random()<0.6 else 400)); seed=int(rng.integers(0,1<<30))
    store=int(rng.integers(0,3)); defer=int(rng.choice([-1,-1,0,6,12,20,32]))
    c=OracleCircuit(n)
    # sparse-ish circuits: restrict to a random subset of qubits half of the time
    if rng.random()<0.5:
        sub=rng.permutation(n)[:int(rng.integers(1,n+1))]
        r2=np.random.default_rng(seed)
        for _ in range(gates):
            k=r2.integers(0,6); t=int(sub[r2.integers(0,len(sub))])
            if k==0: c.add_hadamard_gate(t)
            elif k==1: c.add_rotation_x_gate(t,float(r2.random()*6.28))
            elif k==2: c.add_rotation_y_gate(t,float(r2.random()*6.28))
            elif k==3: c.add_rotation_z_gate(t,float(r2.random()*6.28))
            elif k==4: c.add_pauli_x_gate(t,False)
            else:
                ctl=int(r2.integers(0,n))
                if ctl!=t: c.add_cnot_gate(ctl,t)
    else:
        circuits.random_circuit(c,n,gates,seed)
    got,_=emu_run(c,world,track_support=True,store_side=store,defer=defer)
    got2,_=emu_run(c,world,store_side=store,defer=defer)
    c.forward()
    e1=rel_err(got,c.amplitudes()); e2=rel_err(got2,c.amplitudes())
    if not (e1<1e-12 and e2<1e-12) or np.isnan(got.view(np.float64)).any():
        print('FAIL',n,world,gates,seed,store,defer,e1,e2); sys.exit(1)
    n_ok+=1
print('ok',n_ok,'circuits')
