"""Parent-array skeleton container (host-side input contract of the lifting path).

Mirrors the interface of the reference `common/skeleton.py:4-81` (`Skeleton(parents,
joints_left, joints_right)`, `num_joints()`, `parents()`, `remove_joints()`), written
from its behaviour: a skeleton is a parent index per joint, -1 for the root.
"""
import numpy as np


class Skeleton:
    def __init__(self, parents, joints_left, joints_right):
        assert len(joints_left) == len(joints_right)
        self._parents = parents
        self._joints_left = joints_left
        self._joints_right = joints_right

    def num_joints(self):
        return len(self._parents)

    def parents(self):
        return self._parents

    def joints_left(self):
        return self._joints_left

    def joints_right(self):
        return self._joints_right

    def has_children(self):
        return self._has_children

    def children(self):
        return self._children

    def remove_joints(self, joints_to_remove):
        """Drop joints, re-parent orphans to their nearest kept ancestor, re-index.

        Returns the list of kept (old) joint indices, like the reference
        (`common/skeleton.py:24-62`).
        """
        drop = set(int(j) for j in joints_to_remove)
        n = len(self._parents)
        parents = list(self._parents)
        for i in range(n):
            while parents[i] in drop:
                parents[i] = parents[parents[i]]
        kept = [j for j in range(n) if j not in drop]
        new_index = {old: new for new, old in enumerate(kept)}
        self._parents = np.array(
            [new_index[parents[j]] if parents[j] >= 0 else parents[j] for j in kept])
        if self._joints_left is not None:
            self._joints_left = [new_index[j] for j in self._joints_left if j in new_index]
        if self._joints_right is not None:
            self._joints_right = [new_index[j] for j in self._joints_right if j in new_index]
        self._compute_metadata()
        return kept

    def _compute_metadata(self):
        n = len(self._parents)
        self._has_children = np.zeros(n).astype(bool)
        self._children = [[] for _ in range(n)]
        for i, p in enumerate(self._parents):
            if p != -1:
                self._has_children[p] = True
                self._children[p].append(i)
