"""Extracts the KyteaFullwidthFilter character map from the reference source (run in the build container,
where /root/reference exists) into tests/golden/kytea_fullwidth_map.json: {source code point: target code point}.
The oracle's and the library's own tables are checked against this fixture."""
import json
import os
import re

SRC = "/root/reference/vaporetto_rules/src/string_filters/kytea_fullwidth.rs"
HERE = os.path.dirname(os.path.abspath(__file__))


def unescape(s: str) -> str:
    return {"\\'": "'", "\\\\": "\\"}.get(s, s)


def main():
    text = open(SRC, encoding="utf-8").read()
    pairs = re.findall(r"'((?:\\.|[^'\\]))' => '((?:\\.|[^'\\]))'", text)
    m = {}
    for a, b in pairs:
        a, b = unescape(a), unescape(b)
        assert len(a) == 1 and len(b) == 1, (a, b)
        assert ord(a) not in m
        m[ord(a)] = ord(b)
    with open(os.path.join(HERE, "kytea_fullwidth_map.json"), "w") as f:
        json.dump({str(k): v for k, v in sorted(m.items())}, f, indent=0)
    print(len(m), "entries")


if __name__ == "__main__":
    main()
