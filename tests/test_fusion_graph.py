"""Container semantics of ObjectType / Relation / FusionGraph (host code, no GPU).
Behaviours follow the reference's skfusion/tests/test_fusion_graph.py and fusion_graph.py."""
import numpy as np
import pytest

from skfusion.fusion import FusionGraph, ObjectType, Relation
from skfusion.fusion.graph import DataFusionError


def _graph():
    rs = np.random.RandomState(0)
    t1, t2, t3 = ObjectType('Type 1', 10), ObjectType('Type 2', 20), ObjectType('Type 3', 30)
    rels = [Relation(rs.rand(50, 100), t1, t2, name='r12'), Relation(rs.rand(50, 40), t1, t3),
            Relation(rs.rand(100, 40), t2, t3), Relation(rs.rand(50, 50), t1, t1, name='theta1')]
    return FusionGraph(rels), (t1, t2, t3), rels


def test_object_type_identity_is_by_name():
    a, b = ObjectType('x', 3), ObjectType('x', 7)
    assert a == b and hash(a) == hash(b) == hash('x') and a != ObjectType('y')
    assert repr(a) == 'ObjectType("x")' and str(a) == 'x'
    assert ObjectType('z').rank == 5


def test_relations_compare_by_id_but_hash_by_label():
    t1, t2 = ObjectType('a'), ObjectType('b')
    r1, r2 = Relation(np.zeros((2, 2)), t1, t2), Relation(np.zeros((2, 2)), t1, t2)
    assert hash(r1) == hash(r2) and r1 != r2             # unnamed parallel relations stay distinct
    n1, n2 = Relation(np.zeros((2, 2)), t1, t2, name='same'), Relation(np.ones((2, 2)), t1, t2, name='same')
    assert n1 == n2
    assert t1 in r1 and ObjectType('c') not in r1
    assert str(n1) == 'Relation(a "same" b)' and repr(r1) == 'Relation(ObjectType("a") → ObjectType("b"))'
    extra = Relation(np.zeros((1, 1)), t1, t2, source='paper')
    assert extra.source == 'paper' and extra.fill_value == 'mean'


def test_graph_bookkeeping_and_queries():
    g, (t1, t2, t3), rels = _graph()
    assert g.n_relations == 4 and g.n_object_types == 3
    assert list(g.object_types) == [t1, t2, t3]
    assert list(g.get_relations(t1, t2)) == [rels[0]] and list(g.get_relations(t3, t1)) == []
    assert g.get_relation('r12') is rels[0] and g['r12'] is rels[0]
    assert g.get_object_type('Type 2') is t2
    assert set(g.out_neighbors(t1)) == {t2, t3, t1} and set(g.in_neighbors(t3)) == {t1, t2}
    assert list(g.out_relations(t2)) == [rels[2]] and set(g.in_relations(t3)) == {rels[1], rels[2]}
    with pytest.raises(DataFusionError):
        g.get_relation('nope')
    with pytest.raises(DataFusionError):
        list(g.get_relations(t1, ObjectType('other')))
    assert str(g) == 'FusionGraph(Object types: 3, Relations: 4)'


def test_parallel_relations_keep_insertion_order():
    t1, t2 = ObjectType('a'), ObjectType('b')
    r = [Relation(np.zeros((2, 3)) + i, t1, t2) for i in range(3)]
    g = FusionGraph(r)
    assert [x.data[0, 0] for x in g.get_relations(t1, t2)] == [0, 1, 2]


def test_removing_relations_drops_orphan_types():
    g, (t1, t2, t3), rels = _graph()
    g.remove_relation(rels[2])
    assert g.n_relations == 3 and t2 in g.object_types and t3 in g.object_types
    g.remove_relation(rels[1])
    assert t3 not in g.object_types                         # no relation touches Type 3 any more
    g.remove_relations_from([rels[0]])
    assert t2 not in g.object_types and t1 in g.object_types  # theta keeps Type 1 alive
    g.remove_relation(rels[3])
    assert g.n_object_types == 0 and g.n_relations == 0


def test_remove_object_type_cascades():
    g, (t1, t2, t3), rels = _graph()
    g.remove_object_type(t3)
    assert t3 not in g.object_types and g.n_relations == 2
    assert all(t3 not in r for r in g.relations)


def test_names_and_metadata_merge():
    t1, t2 = ObjectType('a'), ObjectType('b')
    r1 = Relation(np.zeros((2, 3)), t1, t2, row_names=['x', 'y'], row_metadata=[{'p': 1}, {'p': 2}])
    r2 = Relation(np.zeros((3, 2)), t2, t1, col_metadata=[{'q': 5}, {'q': 6}])
    g = FusionGraph([r1, r2])
    assert g.get_names(t1) == ['x', 'y'] and g.get_names('b') == ['0', '1', '2']
    assert g.get_metadata(t1) == [{'p': 1, 'q': 5}, {'p': 2, 'q': 6}]


def test_fill_modes():
    x = np.array([[1., np.nan, 3.], [np.inf, 5., 6.]])
    assert Relation(x, 'a', 'b', fill_value=0).filled()[0, 1] == 0 and Relation(x, 'a', 'b', fill_value=-1.5).filled()[1, 0] == -1.5
    m = np.ma.masked_array([[1., 2., 3.], [4., 5., 6.]], mask=[[0, 1, 0], [0, 0, 0]])
    fm = Relation(m, 'a', 'b').filled()
    # 'mean' / constant fills write the data but the numpy mask survives (that is how Dfmc later
    # recognises the unknown entries, dfmc.py:78-83); row/col means drop it.  Same as upstream.
    assert np.ma.getdata(fm)[0, 1] == pytest.approx(np.mean([1, 3, 4, 5, 6])) and np.ma.is_masked(fm)
    fr = Relation(m, 'a', 'b', fill_value='row_mean').filled()
    assert fr[0, 1] == pytest.approx(2.0) and not np.ma.is_masked(fr)
    fc = Relation(m, 'a', 'b', fill_value='col_mean').filled()
    assert fc[0, 1] == pytest.approx(5.0) and not np.ma.is_masked(fc)
    y = np.array([[1., np.nan], [3., 5.]])
    assert Relation(y, 'a', 'b', fill_value='row_mean').filled()[0, 1] == 1.0
    assert np.isnan(y[0, 1])                                # the caller's data is untouched
