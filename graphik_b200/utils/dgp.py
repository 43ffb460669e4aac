"""Distance-geometry helpers with the reference's names (graphik/utils/dgp.py).

Graph <-> matrix glue runs on the host (it is bookkeeping, built once per
robot); `bound_smoothing` and `MDS`-based initialisation are part of the hot path
and run on the device through the C ABI -- there is no CPU implementation here.
"""
import numpy as np

from graphik_b200.graphs.graph_revolute import GoalGraph


def distance_matrix_from_graph(G: GoalGraph, label=None, nonedge=0) -> np.ndarray:
    """dgp.py:42-50: squared DIST per edge; edges WITHOUT a DIST attribute show up as 1.0
    (networkx's default weight, squared) exactly as in the reference; non-edges are 0."""
    return np.where(np.isnan(G.dist), np.where(G.edge, 1.0, float(nonedge)), G.dist) ** 2


def adjacency_matrix_from_graph(G: GoalGraph, label=None, nodelist=None) -> np.ndarray:
    """dgp.py:53-65: 0/1 mask of the edges that carry a DIST."""
    return (~np.isnan(G.dist)).astype(float)


def gram_from_distance_matrix(D):
    """dgp.py:28-31."""
    n = D.shape[0]
    J = np.identity(n) - (1 / n) * np.ones(D.shape)
    return -0.5 * J @ D @ J


def distance_matrix_from_gram(X):
    """dgp.py:34-35."""
    return (X.diagonal()[:, np.newaxis] + X.diagonal()) - 2 * X


def distance_matrix_from_pos(Y):
    """dgp.py:38-39."""
    return distance_matrix_from_gram(Y @ Y.T)


def pos_from_graph(G: GoalGraph, node_ids=None) -> np.ndarray:
    """dgp.py:68-82."""
    if not node_ids:
        return np.array(G.pos)
    index = {u: k for k, u in enumerate(G.node_ids)}
    return np.array([G.pos[index[u]] for u in node_ids])


def graph_from_pos(P, node_ids=None, dist=True) -> GoalGraph:
    """dgp.py:85-104: complete graph over the given points."""
    P = np.asarray(P, dtype=float)
    n = P.shape[0]
    if not node_ids:
        node_ids = ["p" + str(k) for k in range(n)]
    diff = P[:, None, :] - P[None, :, :]
    d = np.sqrt(np.sum(diff * diff, axis=-1))
    edge = ~np.eye(n, dtype=bool)
    dm = np.where(edge, d, np.nan)
    if not dist:
        edge = np.zeros((n, n), bool)
        dm = np.full((n, n), np.nan)
    return GoalGraph(list(node_ids), edge, dm, dm.copy(), dm.copy(), P)


def bound_smoothing(G: GoalGraph):
    """dgp.py:192-231 on the device (gik_bounds): (lower, upper), unsquared."""
    from graphik_b200.engine import BatchIK
    from graphik_b200.plan import Plan
    N = G.number_of_nodes()
    lower = np.where(G.edge & ~np.isnan(G.lower), G.lower, 0.0)
    upper = np.where(G.edge & ~np.isnan(G.upper), G.upper, np.inf)
    np.fill_diagonal(lower, 0.0)
    np.fill_diagonal(upper, 0.0)
    plan = Plan({
        "n_nodes": N, "term_i": np.zeros(0, np.int32), "term_j": np.zeros(0, np.int32),
        "term_kind": np.zeros(0, np.int32), "term_target": np.zeros(0), "term_goal": np.zeros(0, np.int32),
        "n_goal": 0, "anchor_node": None,
        "bs_lower": np.ascontiguousarray(lower), "bs_upper": np.ascontiguousarray(upper),
    })
    eng = BatchIK(plan=plan)
    lb, ub = eng.bounds(None, B=1)
    return lb[0].cpu().numpy(), ub[0].cpu().numpy()
