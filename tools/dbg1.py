import numpy as np, sys
sys.path.insert(0,'/root/repo')
import george_b200 as george, oracle
from george_b200 import kernels
from george_b200._spec import flatten
np.random.seed(12345)
t = np.sort(np.random.uniform(2000, 2010, 50))
y = 340 + 2 * (t - 2000) + 0.5 * np.sin(2 * np.pi * t) + 0.1 * np.random.randn(50)
k = 66.0 ** 2 * kernels.Matern32Kernel(67.0 ** 2) + kernels.ConstantKernel(log_constant=np.log(0.5))
gp1 = george.GP(k, mean=np.mean(y))
gp2 = george.GP(k, mean=np.mean(y), solver=george.HODLRSolver)
gp1.compute(t); gp2.compute(t)
print(gp1.log_likelihood(y), gp2.log_likelihood(y))
print(gp1.solver.log_determinant, gp2.solver.log_determinant)
K = oracle.value_symmetric(flatten(k), t[:,None]) + 1.25e-12*np.eye(50)
print(np.linalg.slogdet(K), np.linalg.cond(K))
r = y-np.mean(y)
import scipy.linalg as sl
c = sl.cholesky(K, lower=True)
print("scipy ll", -0.5*(50*np.log(2*np.pi)+2*np.sum(np.log(np.diag(c)))) - 0.5*r@sl.cho_solve((c,True), r))
o = oracle.HODLR(flatten(k), t, np.sqrt(1.25e-12)*np.ones(50))
print("oracle ll", -0.5*(50*np.log(2*np.pi)+o.log_determinant) - 0.5*o.dot_solve(r), o.log_determinant)
print(gp1.solver.dot_solve(r), gp2.solver.dot_solve(r), o.dot_solve(r), r@sl.cho_solve((c,True), r))
