"""PhysicalParameters - same attribute surface as the reference class
(thermalporous/physicalparameters.py:3-98).  Values are plain floats; the property laws are
evaluated on the GPU inside libtpb200 (csrc/tpb_internal.cuh), the NumPy versions below exist for
host-side reporting (rates, masses) and follow the reference formulas line by line."""
import numpy as np


class PhysicalParameters:
    ko = 0.15                    # physicalparameters.py:9
    kw = 0.6005638               # :10
    kr = 1.7295772056            # :12
    c_v_w = 4181.3               # :13
    c_v_o = 2093.4               # :14
    c_r = 920.0                  # :15
    rho_r = 2650.0               # :16
    p_inj = 6.895e7 * 1e-6       # :17
    p_prod = 2.7579e7 * 1e-6     # :18
    T_inj = 422.039              # :20
    T_prod = 288.706             # :23
    API = 10.0                   # :24
    p_ref = 4.1369e7 * 1e-6      # :25
    T_ref = (T_inj + T_prod) / 2.0
    g = 9.80665 * 1e-6           # :27
    S_o = 1.0                    # :28
    U = 5.44409e6                # :29
    rate = 1.8e-3                # :30
    p_ct = 1e5 * 1e-6
    lmbda = 2.0
    Sr_o = 0.1
    Sr_w = 0.2
    well_radius = 0.1            # :35

    def oil_rho(self, p, T):
        """:37-46"""
        SG = 141.5 / (self.API + 131.5)
        rho_ref = SG * 999.0
        return rho_ref * np.exp(5.5e-5 * (p * 1e1 - 1.01325)) * np.exp(-2.5e-4 * (T - (15.5556 + 273.15)))

    def oil_mu(self, T):
        """:48-57"""
        Tf = 1.8 * (T - 273.15) + 32.0
        return 1e-3 * (10.0 ** (-0.8021 * self.API + 23.8765) * Tf ** (0.31458 * self.API - 9.21592))

    def water_rho(self, p, T):
        """:69-82"""
        Tc = T - 272.15
        poly = (999.83952 + 16.955176 * Tc - 7.987e-3 * Tc ** 2 - 46.170461e-6 * Tc ** 3
                + 105.56302e-9 * Tc ** 4 - 280.54353e-12 * Tc ** 5)
        return poly * np.exp(3.98854e-4 * (p - 10.2)) / (1 + 16.87985e-3 * Tc)

    def water_mu(self, T):
        """:84-90"""
        Tf = 1.8 * (T - 272.15) + 32
        return 1e-3 * 2.1850 / (-1 + 0.04012 * Tf + 5.1547e-6 * Tf ** 2)

    def rel_perm_o(self, S_o):
        return S_o

    def rel_perm_w(self, S_o):
        return 1.0 - S_o
