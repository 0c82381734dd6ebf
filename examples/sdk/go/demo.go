// Go client demo for the deeprec_b200 model server: POST /v1/models/<name>:predict with a JSON body.
//
//	go run demo.go -url http://127.0.0.1:8500 -model ctr
package main

import (
	"bytes"
	"encoding/json"
	"flag"
	"fmt"
	"io"
	"math/rand"
	"net/http"
	"os"
)

type predictRequest struct {
	Dense [][]float32 `json:"dense"` // [B][13]
	Ids   [][]int64   `json:"ids"`   // [26][B]
}

type predictResponse struct {
	Predictions  []float32 `json:"predictions"`
	ModelVersion int64     `json:"model_version"`
}

func main() {
	url := flag.String("url", "http://127.0.0.1:8500", "model server base URL")
	model := flag.String("model", "ctr", "model name")
	batch := flag.Int("batch", 4, "rows per request")
	flag.Parse()

	req := predictRequest{Dense: make([][]float32, *batch), Ids: make([][]int64, 26)}
	for i := range req.Dense {
		req.Dense[i] = make([]float32, 13)
		for k := range req.Dense[i] {
			req.Dense[i][k] = rand.Float32()*2 - 1
		}
	}
	for t := range req.Ids {
		req.Ids[t] = make([]int64, *batch)
		for i := range req.Ids[t] {
			req.Ids[t][i] = rand.Int63n(1000)
		}
	}
	body, err := json.Marshal(req)
	if err != nil {
		fmt.Fprintln(os.Stderr, err)
		os.Exit(1)
	}
	resp, err := http.Post(fmt.Sprintf("%s/v1/models/%s:predict", *url, *model), "application/json", bytes.NewReader(body))
	if err != nil {
		fmt.Fprintln(os.Stderr, err)
		os.Exit(1)
	}
	defer resp.Body.Close()
	raw, _ := io.ReadAll(resp.Body)
	if resp.StatusCode != http.StatusOK {
		fmt.Fprintf(os.Stderr, "HTTP %d: %s\n", resp.StatusCode, raw)
		os.Exit(1)
	}
	var out predictResponse
	if err := json.Unmarshal(raw, &out); err != nil {
		fmt.Fprintln(os.Stderr, err)
		os.Exit(1)
	}
	fmt.Println("model version", out.ModelVersion, "predictions", out.Predictions)
}
