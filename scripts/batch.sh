#!/bin/bash
# Time-breakdown trio (reference dear/batch.sh:38-42): full DeAR, without the all-gather (FF part),
# without the reduce-scatter (BP part).
here="$(cd "$(dirname "$0")" && pwd)"
dnn="${dnn:-resnet50}"; bs="${bs:-64}"; nworkers="${nworkers:-8}"
mkdir -p "$here/../logs/breakdown"
for parts in "" "allgather" "reducescatter"; do
  tag="${parts:-none}"
  dnn=$dnn bs=$bs nworkers=$nworkers exclude_parts="$parts" "$here/launch.sh" \
    > "$here/../logs/breakdown/${dnn}-bs${bs}-n${nworkers}-exclude-${tag}.log" 2>&1
  grep "Total" "$here/../logs/breakdown/${dnn}-bs${bs}-n${nworkers}-exclude-${tag}.log" | tail -1
done
