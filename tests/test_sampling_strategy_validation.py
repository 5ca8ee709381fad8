"""CPU: gigl_b200.dag.validate_strategy against the cases of the reference's own config-validation tests
(python/tests/unit/src/validation/subgraph_sampling_strategy_validation_test.py:176-650 and
sampling_op_validation_test.py:77-330), on the reference's test assets (python/tests/test_assets/graph_metadata_constants.py:
node types '0' '1' '2', edge types 0->1, 0->2, 1->2, NABLP supervision on 0->1; the homogeneous default 0 -0-> 0).  Same
inputs, same error type per case."""
import pytest

from gigl_b200 import dag

IN, OUT = dag.INCOMING, dag.OUTGOING


def et(src, rel, dst):
    return {"srcNodeType": src, "relation": rel, "dstNodeType": dst}


HOMO_ET = et("0", "0", "0")
HOMO_GRAPH = {"nodeTypes": ["0"], "edgeTypes": [HOMO_ET]}
HOMO_TASK = {"nodeAnchorBasedLinkPredictionTaskMetadata": {"supervisionEdgeTypes": [HOMO_ET]}}
E01, E02, E12 = et("0", "0", "1"), et("0", "1", "2"), et("1", "2", "2")
HET_GRAPH = {"nodeTypes": ["0", "1", "2"], "edgeTypes": [E01, E02, E12]}
HET_TASK = {"nodeAnchorBasedLinkPredictionTaskMetadata": {"supervisionEdgeTypes": [E01]}}


def op(name, edge_type, inputs=(), direction=IN, n=10):
    return {"opName": name, "edgeType": edge_type, "inputOpNames": list(inputs), "randomUniform": {"numNodesToSample": n},
            "samplingDirection": direction}


def path(root, *ops):
    return {"rootNodeType": root, "samplingOps": list(ops)}


def rejected(paths, graph, task, error_type):
    with pytest.raises(dag.SamplingValidationError) as e:
        dag.validate_strategy(paths, graph, task)
    assert e.value.error_type == error_type, e.value


def test_example_dags_validate():
    # :176-188 - the homogeneous example and the heterogeneous one (root '0' OUTGOING over 0->1, root '1' INCOMING over 0->1)
    dag.validate_strategy([path("0", op("example_homogeneous_sampling_op", HOMO_ET))], HOMO_GRAPH, HOMO_TASK)
    got = dag.validate_strategy([path("0", op("example_heterogeneous_sampling_op_0", E01, direction=OUT)),
                                 path("1", op("example_heterogeneous_sampling_op_1", E01))], HET_GRAPH, HET_TASK)
    assert sorted(got) == ["0", "1"] and got["0"][0].sampling_direction == OUT


def test_repeated_root_node_type():  # :190-220
    p = path("0", op("example_homogeneous_sampling_op", HOMO_ET))
    rejected([p, p], HOMO_GRAPH, HOMO_TASK, "REPEATED_ROOT_NODE_TYPE")


def test_repeated_op_name_is_per_path():  # :222-307
    rejected([path("2", op("locally_repeated_sampling_op", E02), op("locally_repeated_sampling_op", E12))], HET_GRAPH, HET_TASK,
             "REPEATED_OP_NAME")
    # the same name in two different paths is fine at construction (the reference builds it without validating further)
    for p in ([path("1", op("globally_repeated_sampling_op", E01))], [path("2", op("globally_repeated_sampling_op", E02))]):
        assert dag.plan(dag.ops_from_config(p[0]), p[0]["rootNodeType"])


def test_bad_input_op_name():  # :309-341
    rejected([path("1", op("nonexistent_input_name_sampling_op", E01, ["nonexistent_input_name"]))], HET_GRAPH, HET_TASK, "BAD_INPUT_OP_NAME")


def test_edge_type_not_in_graph_metadata():  # :343-374
    rejected([path("0", op("nonexistent_edge_type_sampling_op", E01, direction=OUT))], HOMO_GRAPH, HOMO_TASK,
             "SAMPLING_OP_EDGE_TYPE_NOT_IN_GRAPH_METADATA")


def test_cycle_detection():  # :376-485
    cyc = [op("cycle_op_1", HOMO_ET, direction=OUT), op("cycle_op_2", HOMO_ET, ["cycle_op_1", "cycle_op_4"], OUT),
           op("cycle_op_3", HOMO_ET, ["cycle_op_2"], OUT), op("cycle_op_4", HOMO_ET, ["cycle_op_3"], OUT)]
    rejected([path("0", *cyc)], HOMO_GRAPH, HOMO_TASK, "DAG_CONTAINS_CYCLE")
    ok = [op("no_cycle_op_1", HOMO_ET, direction=OUT), op("no_cycle_op_2", HOMO_ET, ["no_cycle_op_1"], OUT),
          op("no_cycle_op_3", HOMO_ET, ["no_cycle_op_2", "no_cycle_op_1"], OUT)]
    got = dag.validate_strategy([path("0", *ok)], HOMO_GRAPH, HOMO_TASK)
    # an op with two inputs runs once per input (GraphDBSampler.scala:66-82 expands the union)
    assert [p.key for p in dag.plan(got["0"], "0")] == ["no_cycle_op_1", "no_cycle_op_2", "no_cycle_op_3@no_cycle_op_2", "no_cycle_op_3@no_cycle_op_1"]


def test_root_node_type_not_in_graph_metadata():  # :487-514
    rejected([path("2", op("example_homogeneous_sampling_op", HOMO_ET))], HOMO_GRAPH, HOMO_TASK, "ROOT_NODE_TYPE_NOT_IN_GRAPH_METADATA")


def test_root_node_type_not_in_task_metadata():  # :516-554 - supervision is 0->1; a path rooted at '2' has no business here
    rejected([path("0", op("example_heterogeneous_sampling_op_0", E01, direction=OUT)), path("1", op("example_heterogeneous_sampling_op_1", E01)),
              path("2", op("example_heterogeneous_sampling_op_2", E02))], HET_GRAPH, HET_TASK, "ROOT_NODE_TYPE_NOT_IN_TASK_METADATA")


def test_expected_root_node_type_missing():  # :556-582 - only '0' given, the supervision edge also roots samples at '1'
    rejected([path("0", op("example_heterogeneous_sampling_op_0", E01, direction=OUT))], HET_GRAPH, HET_TASK, "MISSING_EXPECTED_ROOT_NODE_TYPE")


def test_no_root_sampling_op():  # :584-625
    rejected([path("0", op("no_root_sampling_op_1", HOMO_ET, ["no_root_sampling_op_2"], OUT),
                   op("no_root_sampling_op_2", HOMO_ET, ["no_root_sampling_op_1"], OUT))], HOMO_GRAPH, HOMO_TASK, "MISSING_ROOT_SAMPLING_OP")


def test_zero_hop_dag_is_valid():  # :627-650
    assert dag.validate_strategy([path("0")], HOMO_GRAPH, HOMO_TASK) == {"0": []}
    assert dag.plan([], "0") == []


# ---- sampling_op_validation_test.py: the edge-type rule of one op against the root / its parent ---------------------
def test_root_op_direction_rules():  # :77-128
    rejected([path("0", op("incoming_root_sampling_op", E01))], HET_GRAPH, {"nodeBasedTaskMetadata": {"supervisionNodeTypes": ["0"]}},
             "CONTAINS_INVALID_EDGE_IN_DAG")                                     # INCOMING over 0->1 expands '1' nodes
    dag.validate_strategy([path("1", op("incoming_root_sampling_op", E01))], HET_GRAPH, {"nodeBasedTaskMetadata": {"supervisionNodeTypes": ["1"]}})
    rejected([path("1", op("outgoing_root_sampling_op", E01, direction=OUT))], HET_GRAPH, {"nodeBasedTaskMetadata": {"supervisionNodeTypes": ["1"]}},
             "CONTAINS_INVALID_EDGE_IN_DAG")                                     # OUTGOING over 0->1 expands '0' nodes
    dag.validate_strategy([path("0", op("outgoing_root_sampling_op", E01, direction=OUT))], HET_GRAPH,
                          {"nodeBasedTaskMetadata": {"supervisionNodeTypes": ["0"]}})


@pytest.mark.parametrize("root,parent,child,ok", [
    # child INCOMING, parent INCOMING (:130-190): child.dst == parent.src
    ("2", ("p", E12, IN), ("c", E01, IN), True), ("2", ("p", E02, IN), ("c", E01, IN), False),
    # child INCOMING, parent OUTGOING (:196-261): child.dst == parent.dst
    ("1", ("p", E12, OUT), ("c", E02, IN), True), ("0", ("p", E02, OUT), ("c", E01, IN), False),
    # child OUTGOING, parent INCOMING (:267-330): child.src == parent.src
    ("2", ("p", E02, IN), ("c", E01, OUT), True), ("2", ("p", E12, IN), ("c", E02, OUT), False),
    # child OUTGOING, parent OUTGOING: child.src == parent.dst
    ("0", ("p", E01, OUT), ("c", E12, OUT), True), ("0", ("p", E02, OUT), ("c", E01, OUT), False)])
def test_parent_child_direction_rules(root, parent, child, ok):
    task = {"nodeBasedTaskMetadata": {"supervisionNodeTypes": [root]}}
    paths = [path(root, op(parent[0], parent[1], direction=parent[2]), op(child[0], child[1], [parent[0]], child[2]))]
    if ok:
        dag.validate_strategy(paths, HET_GRAPH, task)
    else:
        rejected(paths, HET_GRAPH, task, "CONTAINS_INVALID_EDGE_IN_DAG")


def test_component_rejects_an_invalid_strategy_before_reading_any_data(tmp_path):
    """The typed sampler component runs the same validation first (no input file is opened: none exists here)."""
    import yaml

    from gigl_b200 import subgraph_sampler

    cfg = {"graphMetadata": {"condensedEdgeTypeMap": {"0": E01}, "condensedNodeTypeMap": {"0": "0", "1": "1"}, "edgeTypes": [E01, E02, E12],
                             "nodeTypes": ["0", "1", "2"]},
           "taskMetadata": HET_TASK,
           "datasetConfig": {"subgraphSamplerConfig": {"numPositiveSamples": 1, "subgraphSamplingStrategy": {"messagePassingPaths": {"paths": [
               path("0", op("a", E01, direction=OUT))]}}}},
           "sharedConfig": {"isGraphDirected": True, "preprocessedMetadataUri": "preprocessed_metadata.yaml",
                            "flattenedGraphMetadata": {"nodeAnchorBasedLinkPredictionOutput": {"tfrecordUriPrefix": "out/", "nodeTypeToRandomNegativeTfrecordUriPrefix": {}}}}}
    (tmp_path / "preprocessed_metadata.yaml").write_text(yaml.safe_dump({"condensedNodeTypeToPreprocessedMetadata": {"0": {}, "1": {}},
                                                                        "condensedEdgeTypeToPreprocessedMetadata": {"0": {}}}))
    (tmp_path / "cfg.yaml").write_text(yaml.safe_dump(cfg))
    with pytest.raises(dag.SamplingValidationError) as e:
        subgraph_sampler.run("cfg.yaml", "job", None, root=str(tmp_path), log=lambda *_: None)
    assert e.value.error_type == "MISSING_EXPECTED_ROOT_NODE_TYPE"
