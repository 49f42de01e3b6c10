"""Physical constants and run parameters of the hot path (same values as py/SaclayMocks/constant.py)."""
c = 299792.458          # km/s
lya = 1215.67
lylimit = 0.0
lyb = 1025.72
lambda_min = 3476.0
h = 0.6731
omega_M_0 = 0.31457
omega_lambda_0 = 0.68543
omega_k_0 = 0.0
QSO_bias = 3.7
z_QSO_bias_1 = 1.9
z_QSO_bias_2 = 2.75
z_QSO_bias_3 = 3.6
z0 = 1.70975268202
H0 = 100.0
