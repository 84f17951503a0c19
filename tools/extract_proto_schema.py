"""Extract the wire contract of the reference's plan protobuf (message -> field name -> [number, type, repeated]; enum -> name ->
value) into tests/golden/auron_proto_schema.json.  Run in the build container (the reference is not on the GPU box):

    python tools/extract_proto_schema.py /root/reference/native-engine/auron-planner/proto/auron.proto

Only numbers, names and types are recorded: it is the table a protobuf decoder needs, which tests/test_proto_contract.py uses to
check that every plan auron_b200/proto.py encodes (the plans all GPU parity tests run) is a well-formed message of the reference."""
import json
import os
import re
import sys


def parse(text: str) -> dict:
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"//[^\n]*", "", text)
    out = {"messages": {}, "enums": {}}
    for kind, name, body in re.findall(r"\b(message|enum)\s+(\w+)\s*\{((?:[^{}]|\{[^{}]*\})*)\}", text):
        if kind == "enum":
            out["enums"][name] = {k: int(v) for k, v in re.findall(r"(\w+)\s*=\s*(-?\d+)\s*;", body)}
            continue
        body = re.sub(r"\boneof\s+\w+\s*\{", "", body).replace("}", "")
        fields = {}
        for rep, typ, fname, num in re.findall(r"(repeated\s+|optional\s+)?([\w.]+)\s+(\w+)\s*=\s*(\d+)\s*(?:\[[^\]]*\])?\s*;", body):
            fields[fname] = [int(num), typ, rep.strip() == "repeated"]
        out["messages"][name] = fields
    return out


if __name__ == "__main__":
    src = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/native-engine/auron-planner/proto/auron.proto"
    schema = parse(open(src).read())
    dst = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "auron_proto_schema.json")
    with open(dst, "w") as f:
        json.dump(schema, f, indent=1, sort_keys=True)
    print(f"{len(schema['messages'])} messages, {len(schema['enums'])} enums -> {dst}")
