"""Test plants, restated from reference tests/src/SimModels.h (test fixtures, numpy)."""
import numpy as np

G = 9.80665


class CentroidalSim:
    """reference tests/src/SimModels.h:233-332: 18-state linear plant, ZOH-discretised.

    state = [c(3), euler(3), v(3), omega(3), P(3), L(3)], input = (force, moment about CoM).
    A has only the block d(pos)/dt = vel, so A^2 = 0 and the ZOH matrices are exact polynomials
    (SURVEY.md App. D): Ad = I + A dt, Bd = (I dt + A dt^2/2) B, Ed = (I dt + A dt^2/2) E.
    """

    def __init__(self, mass, inertia, sim_dt):
        A = np.zeros((18, 18))
        A[0:6, 6:12] = np.eye(6)
        B = np.zeros((18, 6))
        B[6:9, 0:3] = np.eye(3) / mass
        B[9:12, 3:6] = np.diag(1.0 / np.asarray(inertia))
        B[12:18, 0:6] = np.eye(6)
        E = np.zeros(18)
        E[8] = -G
        E[14] = -mass * G
        M = np.eye(18) * sim_dt + A * (sim_dt**2 / 2)
        self.Ad, self.Bd, self.Ed = np.eye(18) + A * sim_dt, M @ B, M @ E
        self.x = np.zeros(18)
        self.mass = mass

    pos = property(lambda s: s.x[0:3])
    vel = property(lambda s: s.x[6:9])
    angular_momentum = property(lambda s: s.x[15:18])

    def update(self, force, moment):
        self.x = self.Ad @ self.x + self.Bd @ np.concatenate([force, moment]) + self.Ed

    def add_disturb(self, lin_impulse_per_mass, ang_impulse_per_mass):
        self.x[6:9] += lin_impulse_per_mass
        self.x[9:12] += ang_impulse_per_mass


class ComZmpSim2d:
    """reference tests/src/SimModels.h:11-138: x'' = omega^2 (x - zmp) per axis, exact ZOH
    (Ad = [[ch, sh/w], [w sh, ch]], Bd = (1 - ch, -w sh))."""

    def __init__(self, com_height, sim_dt):
        w = np.sqrt(G / com_height)
        ch, sh = np.cosh(w * sim_dt), np.sinh(w * sim_dt)
        self.Ad = np.array([[ch, sh / w], [w * sh, ch]])
        self.Bd = np.array([1 - ch, -w * sh])
        self.x, self.y = np.zeros(2), np.zeros(2)

    pos = property(lambda s: np.array([s.x[0], s.y[0]]))
    vel = property(lambda s: np.array([s.x[1], s.y[1]]))

    def update(self, zmp):
        self.x = self.Ad @ self.x + self.Bd * zmp[0]
        self.y = self.Ad @ self.y + self.Bd * zmp[1]

    def add_disturb(self, impulse_per_mass):
        # the reference adds impulse.x() to both axes (SimModels.h:127-128)
        self.x[1] += impulse_per_mass[0]
        self.y[1] += impulse_per_mass[0]


class ComZmpSim3d:
    """reference tests/src/SimModels.h:140-226: per tick the horizontal model is rebuilt with the current
    CoM height; the vertical model is a double integrator with gravity (exact ZOH)."""

    def __init__(self, mass, sim_dt):
        self.mass, self.dt = mass, sim_dt
        self.x, self.y, self.z = np.zeros(2), np.zeros(2), np.zeros(2)

    pos = property(lambda s: np.array([s.x[0], s.y[0], s.z[0]]))
    vel = property(lambda s: np.array([s.x[1], s.y[1], s.z[1]]))

    def update(self, zmp, force_z):
        w = np.sqrt(G / self.z[0])
        ch, sh = np.cosh(w * self.dt), np.sinh(w * self.dt)
        Ad, Bd = np.array([[ch, sh / w], [w * sh, ch]]), np.array([1 - ch, -w * sh])
        self.x = Ad @ self.x + Bd * zmp[0]
        self.y = Ad @ self.y + Bd * zmp[1]
        a = force_z / self.mass - G
        self.z = np.array([self.z[0] + self.dt * self.z[1] + 0.5 * self.dt**2 * a, self.z[1] + self.dt * a])

    def add_disturb(self, impulse_per_mass):
        self.x[1] += impulse_per_mass[0]
        self.y[1] += impulse_per_mass[0]
